#!/bin/bash
# Turn what tools/gpu_round.sh <tag> brought back in gpurun_out/ into the tracked summaries under profiles/ (no GPU needed).
tag=${1:?tag}
LIB=cognitive-radio-network_b200/libcrnsense.so
for n in 1024 8192 2048; do python tools/ncu_summary.py gpurun_out/${tag}_sense_n$n.ncu-rep profiles/${tag}_ncu_sense_n$n > /dev/null; done
python tools/launch_summary.py gpurun_out/${tag}_launches_bench.csv > profiles/${tag}_launches_bench.md
cp gpurun_out/${tag}_launches_bench.csv profiles/
for f in bench_n1 bench_reference bench_wideband bench_multiradio bench_refexact bench_sc16 many_radios_ref; do cp gpurun_out/${tag}_$f.json profiles/; done
cp gpurun_out/${tag}_latency.txt profiles/
cp gpurun_out/${tag}_sweep.json profiles/${tag}_sweep.txt
{ echo "# Frame-loop opcode histograms (tools/sass_loop.py on libcrnsense.so as committed; static SASS, no GPU needed)"; echo
  echo "Packed FP32 (\`FFMA2/FADD2/FMUL2\`) occupies the FP32 pipe for two cycles per warp instruction.  \`LDTM\` / \`STTM\` are the tensor-memory reads / writes (tcgen05.ld / tcgen05.st: twiddle rows, window pairs, all-bins accumulators) of the hybrid plans ($(cuobjdump -sass $LIB | grep -c LDTM) \`LDTM\`, $(cuobjdump -sass $LIB | grep -c STTM) \`STTM\`, $(cuobjdump -sass $LIB | grep -c UTCATOMSWS) allocation instructions in the library; the bulk-copy staging of round 2 is no longer compiled in: $(cuobjdump -sass $LIB | grep -c UBLKCP) \`UBLKCP\`)."; echo
  for pat in 'PlanILi1024ELi32ELi32ELi32ELi1ELi4ELi4EEELb1ELi1ELi0ELb0ELj2148284473E' 'PlanILi1024ELi32ELi32ELi32ELi1ELi4ELi4EEELb1ELi1ELi0ELb0ELj4294967295E' 'HybridPlanILi2048ELi4ELi2EEELb1ELi1ELi0ELb0ELj2148284473E' 'HybridPlanILi4096ELi2ELi2EEELb1ELi1ELi0ELb0ELj2148284473E' 'HybridPlanILi4096ELi2ELi2EEELb1ELi1ELi0ELb0ELj4294967295E' 'HybridPlanILi8192ELi2ELi1EEELb1ELi1ELi0ELb0ELj4294967295E' 'HybridPlanILi8192ELi2ELi1EEELb1ELi1ELi0ELb0ELj2148284473E' 'PlanILi512ELi32ELi32ELi16ELi1ELi8ELi4EEELb0ELi0ELi1ELb0ELj2148284473E'; do python tools/sass_loop.py $LIB "$pat" --md; done; } > profiles/${tag}_sass_frame_loops.md
for k in 1024 2048 8192; do ncu -i gpurun_out/${tag}_sense_n$k.ncu-rep --page source --csv --print-source sass > /tmp/${tag}_src_$k.csv 2>/dev/null; done
python - "$tag" <<'PY'
import csv, collections, json, sys
tag = sys.argv[1]
out = ["# Warp-state sampling of the three profiled kernels (ncu --set full, source page; profiles/%s_ncu_sense_n*.md hold the counters)" % tag, "",
       "Share of all warp samples per stall reason - what a resident warp is doing when it is not issuing (`selected` = issuing).", "",
       "| reason | N=1024 configs[1] | N=2048 configs[3] | N=8192 configs[2] |", "|---|---|---|---|"]
tabs = {}
for k in (1024, 2048, 8192):
    rows = list(csv.reader(open('/tmp/%s_src_%d.csv' % (tag, k))))
    hdr = rows[1]
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    ismp = hdr.index('# Samples'); iw = hdr.index('L1 Wavefronts Shared')
    n = 0; wf = 0.0; tot = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        n += int(r[ismp]); wf += float(r[iw] or 0)
        for i in cols: tot[hdr[i]] += int(r[i])
    tabs[k] = (n, tot, wf)
for key, _ in tabs[8192][1].most_common(10):
    out.append("| %s | %s |" % (key.replace('stall_', ''), " | ".join("%.1f %%" % (100 * tabs[k][1][key] / tabs[k][0]) for k in (1024, 2048, 8192))))
out += ["", "Shared-memory wavefronts per launch (sum over instructions): " + ", ".join("N=%d: %.1f M" % (k, tabs[k][2] / 1e6) for k in (1024, 2048, 8192)), ""]
open('profiles/%s_stall_breakdown.md' % tag, 'w').write("\n".join(out))
t = json.load(open('profiles/traffic.json'))
names = {'n1024': ('sense_n1024_r32x32x1_hann_magsq_cta_refbins', 7999586304, 'BASELINE configs[1], one launch'),
         'n8192': ('sense_n8192_r8x32x32_hann_magsq_cta', 7998537728, 'BASELINE configs[2], one launch'),
         'n2048': ('sense_n2048_r2x32x32_hann_magsq_cta_refbins', 4096 * 64 * 2048 * 8, 'BASELINE configs[3], 4096 streams, one launch')}
unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
for k, (name, alg, what) in names.items():
    m = json.load(open('profiles/%s_ncu_sense_%s.json' % (tag, k)))[0]
    rd = m['dram__bytes_read.sum'] * unit[m['dram__bytes_read.sum.unit']]
    wr = m['dram__bytes_write.sum'] * unit[m['dram__bytes_write.sum.unit']]
    t['kernels'][name] = {'dram_bytes_per_launch': rd + wr, 'algorithmic_bytes_per_launch': alg, 'ratio': (rd + wr) / alg,
                          'source': 'profiles/%s_ncu_sense_%s.md (ncu --set full, %s)' % (tag, k, what)}
    print(name, 'traffic ratio %.4f' % ((rd + wr) / alg), m['gpu__time_duration.sum'], m['gpu__time_duration.sum.unit'])
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
PY

"""C-ABI: the library loads, exports every symbol include/crnsense.h declares, struct layouts match the
ctypes mirrors, and the host-only entry points behave (no GPU compute here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crnsense.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crn_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(crn):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(crn.lib, s), "libcrnsense.so does not export %s" % s
    # and the python mirror binds exactly the declared set
    assert sorted(crn.API) == syms


def test_struct_layouts_match_the_header(crn, tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "crnsense.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(crn_config), sizeof(crn_result), sizeof(crn_synth_config), sizeof(crn_kernel_info));
  printf("%zu %zu %zu %zu ", offsetof(crn_config, segs), offsetof(crn_config, ann_wih), offsetof(crn_config, ann_threshold), offsetof(crn_config, device));
  printf("%zu %zu %zu\n", offsetof(crn_result, ann_out), offsetof(crn_result, feat), offsetof(crn_synth_config, hop_mode));
  return 0;
}''')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(crn.Config), C.sizeof(crn.Result), C.sizeof(crn.SynthConfig), C.sizeof(crn.KernelInfo),
            crn.Config.segs.offset, crn.Config.ann_wih.offset, crn.Config.ann_threshold.offset,
            crn.Config.device.offset, crn.Result.ann_out.offset, crn.Result.feat.offset,
            crn.SynthConfig.hop_mode.offset]
    assert got == want


def test_reference_config_matches_the_reference_literals(crn):
    """CE_Predictive_Node.hpp:31-32, .cpp:173-191 (bins), .cpp:78-120 (weights), .cpp:245 (0.8)."""
    c = crn.config_reference()
    assert (c.nfft, c.frame_len, c.navg) == (512, 512, 10)
    assert (c.window, c.detector, c.postop, c.decide) == (crn.WINDOW_RECT, crn.DET_MAG, crn.POST_SQUARE_OF_SUM, crn.DECIDE_ANN)
    segs = [(c.segs[i].band, c.segs[i].lo, c.segs[i].hi) for i in range(c.nsegs)]
    assert segs == [(0, 300, 310), (1, 0, 16), (1, 496, 511), (2, 55, 85), (3, 189, 222)]
    assert sum(hi - lo for b, lo, hi in segs if b == 1) == 31  # bin 511 excluded upstream
    assert c.ann_wih[0][1] == -0.188208 and c.ann_wih[4][5] == 0.609384 and c.ann_wih[2][2] == 0.741944
    assert c.ann_who[0][1] == -7.033320 and c.ann_who[5][3] == -2.552555 and c.ann_who[3][2] == -13.375309
    assert c.ann_threshold == 0.8
    assert crn.validate(c) == crn.OK


def test_welch_and_wideband_configs(crn):
    c = crn.config_welch(1024, 64)
    assert (c.nfft, c.navg, c.window, c.detector, c.postop) == (1024, 64, crn.WINDOW_HANN, crn.DET_MAGSQ, crn.POST_SUM)
    assert [(c.segs[i].lo, c.segs[i].hi) for i in range(c.nsegs)] == [(600, 620), (0, 32), (992, 1022), (110, 170), (378, 444)]
    w = crn.config_wideband(8192, 64, 64)
    assert w.nbands == 64 and w.nsegs == 64 and w.decide == crn.DECIDE_ENERGY
    assert (w.segs[63].lo, w.segs[63].hi) == (63 * 128, 8192)
    assert crn.validate(c) == crn.OK and crn.validate(w) == crn.OK


@pytest.mark.parametrize("mut,status", [
    (lambda c: setattr(c, "nfft", 1000), -1),        # not a power of two
    (lambda c: setattr(c, "nfft", 16384), -7),       # unsupported size
    (lambda c: setattr(c, "frame_len", 513), -1),    # L > N: the reference smashes buffer[512]; we refuse
    (lambda c: setattr(c, "frame_len", 0), -1),
    (lambda c: setattr(c, "frame_stride", 100), -1),
    (lambda c: setattr(c, "navg", 0), -1),
    (lambda c: setattr(c, "nbands", 65), -1),
    (lambda c: setattr(c, "nbands", 3), -1),         # ANN needs NF + 3 channels
    (lambda c: setattr(c, "window", 7), -1),
    (lambda c: setattr(c.segs[0], "hi", 600), -1),
    (lambda c: setattr(c.segs[0], "band", 9), -1),
])
def test_validate_rejects(crn, mut, status):
    c = crn.config_reference()
    mut(c)
    assert crn.validate(c) == status
    assert crn.lib.crn_last_error()  # a message was recorded, nothing exit()ed


def test_error_strings_and_version(crn):
    assert crn.lib.crn_strerror(0) == b"ok"
    for s in range(-7, 0):
        assert crn.lib.crn_strerror(s) not in (b"ok", b"unknown status")
    ma, mi = C.c_int32(-1), C.c_int32(-1)
    assert crn.lib.crn_version(C.byref(ma), C.byref(mi)) == 0 and ma.value == 0 and mi.value >= 1


def test_no_cpu_fallback(crn):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(crn.CrnError) as ei:
        crn.Sensor(crn.config_reference())
    assert ei.value.status == crn.ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may mention it."""
    pkg = os.path.join(ROOT, "cognitive-radio-network_b200")
    for dp, _, fns in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                for line in txt.splitlines():
                    s = line.strip()
                    if s.startswith(("//", "*", "/*", "#", '"""')) or "oracle" not in s.lower():
                        continue
                    assert not re.search(r"(import|include|dlopen|CDLL|open)\b.*oracle", s), (fn, s)


def test_shard_groups_partition(crn):
    for n in (0, 1, 7, 15258, 4096):
        for w in (1, 2, 4, 8):
            parts = [crn.shard_groups(n, w, r) for r in range(w)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for f, c in parts:
                assert f == pos
                pos += c
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_library_is_sm100a_with_packed_fp32_and_tensor_memory():
    """What the shipped binary is made of, checked on its SASS (no GPU needed): sm_100a cubins only; the FFT codelets
    are packed FP32 (FFMA2); the N >= 1024 plans keep their tables in tensor memory (tcgen05.alloc / st / ld show up as
    UTCATOMSWS / STTM / LDTM) and the N = 256 / 512 plans do not (two half-warp teams share a warp there and
    tcgen05 is warp-wide); the frame loops of the headline kernels have no local-memory (spill) traffic."""
    import shutil
    import sys
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "cognitive-radio-network_b200", "libcrnsense.so")
    elfs = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True, check=True).stdout
    cubins = re.findall(r"ELF file\s+\d+:\s+(\S+)", elfs)
    assert cubins and all(c.endswith(".sm_100a.cubin") for c in cubins), cubins
    sys.path[:0] = [os.path.join(ROOT, "tools")]
    import sass_loop
    seen = {}
    all_kernels = list(sass_loop.kernels(lib))
    for name, ins in all_kernels:
        m = re.search(r"sense_kernelINS_(?:10Hybrid)?(?:4)?PlanILi(\d+)E", name)
        if not m:
            continue
        ops = [sass_loop.opname(t) for _, t in ins]
        n = int(m.group(1))
        d = seen.setdefault(n, dict(kernels=0, ffma2=0, ldtm=0, alloc=0))
        d["kernels"] += 1
        d["ffma2"] += ops.count("FFMA2")
        d["ldtm"] += ops.count("LDTM")
        d["alloc"] += ops.count("UTCATOMSWS")
    assert sorted(seen) == [256, 512, 1024, 2048, 4096, 8192], sorted(seen)
    for n, d in seen.items():
        assert d["ffma2"] > 100 * d["kernels"], (n, d)
        if n >= 1024:
            assert d["ldtm"] > 0 and d["alloc"] >= d["kernels"], (n, d)
        else:
            assert d["ldtm"] == 0 and d["alloc"] == 0, (n, d)
    # frame loops of the BASELINE kernels: configs[1] (1024, reference bands) and configs[2] (8192, all bins)
    for pat in ("PlanILi1024ELi32ELi32ELi32ELi1ELi4ELi4EEELb1ELi1ELi0ELb0ELj2148284473E",
                "HybridPlanILi8192ELi2ELi1EEELb1ELi1ELi0ELb0ELj4294967295E"):
        hits = [(nm, ins) for nm, ins in all_kernels if pat in nm]
        assert len(hits) == 1, pat
        body = [sass_loop.opname(t) for _, t in sass_loop.frame_loop(hits[0][1])]
        assert not any(o.startswith(("LDL", "STL")) for o in body), (pat, [o for o in body if o.startswith(("LDL", "STL"))])
        assert body.count("FFMA2") > 300

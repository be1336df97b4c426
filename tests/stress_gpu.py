"""Run-to-run determinism and sampled parity at scale (GPU box; not collected by pytest: python tests/stress_gpu.py).
Lives under tests/ because it uses the oracle as its checker."""
import sys; import os; _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [_R, os.path.join(_R, 'oracle'), os.path.join(_R, 'tests')]
import numpy as np, torch, crn_b200 as crn, oracle
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream
# (nfft, K, groups): ragged K -> barrier-free unit epilogue; K = 64 with few groups -> groups split over CTAs
for (nfft, K, ng) in ((512, 7, 40000), (1024, 5, 40000), (2048, 10, 40000), (512, 3, 40000), (256, 9, 40000),
                      (1024, 64, 200), (8192, 64, 30), (2048, 64, 333), (1024, 64, 1)):
    cfg = crn.config_welch(nfft, K) if nfft >= 512 else crn.config_wideband(nfft, K, 16)
    gs = cfg.group_samples
    d_iq = torch.empty(ng * gs, 2, dtype=torch.float32, device='cuda')
    crn.synth_generate(crn.synth_config(gs, dwell_groups=3, seed=nfft+K), d_iq, 0, ng*gs, None, 0, stream)
    ref = None
    with crn.Sensor(cfg) as s:
        print(s.kernel_info()['name'], 'K', K)
        for rep in range(6):
            d_feat = torch.zeros(ng, cfg.nbands, dtype=torch.float32, device='cuda')
            d_ann = torch.zeros(ng, 3, dtype=torch.float64, device='cuda')
            d_dec = torch.full((ng,), -5, dtype=torch.int32, device='cuda')
            s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, None, stream)
            torch.cuda.synchronize()
            cur = (d_feat.cpu().numpy(), d_ann.cpu().numpy(), d_dec.cpu().numpy())
            if ref is None: ref = cur
            else:
                for a, b in zip(ref, cur): assert np.array_equal(a, b), 'run-to-run mismatch'
    pick = np.random.default_rng(0).integers(0, ng, min(300, ng))
    iq = np.concatenate([d_iq[g*gs:(g+1)*gs].cpu().numpy().view(np.complex64).ravel() for g in pick])
    of, oa, od, _ = oracle.sense_port(cfg, iq, nthreads=8)
    rel = np.abs(ref[0][pick]-of)/np.abs(of)
    print('  max rel', rel.max(), 'dec equal', np.array_equal(ref[2][pick], od) if cfg.decide==1 else 'n/a')
    assert rel.max() < 1e-4
print('stress ok')

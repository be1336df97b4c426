"""The occupancy predictor on its own: batched forward pass (crn_ann_forward_device) and on-device retraining
(crn_ann_train_device) against their CPU statements in oracle/crn_oracle.c.

Pins for the forward pass: the reference engine's own Output[1..3] in tests/golden/ref_*.npz (produced by the
unmodified CE_Predictive_Node.cpp, oracle O1) and the known answers that follow from its 43 weight literals
(SURVEY 8a).  The reference contains no training code (only its outcome, .cpp:74), so the trainer is pinned by a
finite-difference check of the gradient and by the serial CPU statement of the same update rule.

Tolerances: forward outputs 1e-12 absolute (fp64, CUDA exp vs glibc exp differ by <= 1 ulp); trained weights
1e-9 relative and the error 1e-11 relative after a few hundred epochs (the GPU adds the per-example gradients in
a tree and contracts to FMA, the CPU adds in order; measured agreement on the B200: <= 1e-15)."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def ref_weights(crn):
    return crn.AnnWeights.from_config(crn.config_reference())


def labelled_features(rng, n):
    """Feature rows shaped like the engine's (NF^2, CH1, CH2, CH3): the labelled channel ~30 dB above the others."""
    labels = rng.integers(0, 4, n).astype(np.int32)          # 0 = nothing occupied, 1..3 = channel
    feat = (1e5 * rng.uniform(0.5, 2.0, (n, 4))).astype(np.float32)
    feat[:, 0] = (1e4 * rng.uniform(0.5, 2.0, n)).astype(np.float32)
    for k in (1, 2, 3):
        feat[labels == k, k] = (1e8 * rng.uniform(0.5, 2.0, (labels == k).sum())).astype(np.float32)
    return feat, labels


SCALE = [1e-4, 1e-8, 1e-8, 1e-8]


# ---------------------------------------------------------------- CPU: the oracle itself ----------------------

def test_struct_layouts(crn, tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "crnsense.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(crn_ann_weights), offsetof(crn_ann_weights, who), sizeof(crn_ann_train_config),
         offsetof(crn_ann_train_config, input_scale), offsetof(crn_ann_train_config, init_range), offsetof(crn_ann_train_config, seed));
  return 0;
}''')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(crn.AnnWeights), crn.AnnWeights.who.offset, C.sizeof(crn.AnnTrainConfig),
                   crn.AnnTrainConfig.input_scale.offset, crn.AnnTrainConfig.init_range.offset,
                   crn.AnnTrainConfig.seed.offset]
    # the weight block is laid out exactly like crn_config.ann_wih / ann_who
    assert crn.Config.ann_who.offset - crn.Config.ann_wih.offset == crn.AnnWeights.who.offset


def test_oracle_forward_known_answers(crn, oracle):
    """what the reference's weight literals (.cpp:78-120) give for zero and saturating inputs."""
    w = ref_weights(crn)
    out, dec = oracle.ann_forward(w, np.zeros((1, 4), np.float32))
    assert np.allclose(out[0], [4.78996574e-01, 4.12229629e-05, 3.35047425e-03], rtol=1e-8)
    assert dec[0] == crn.ALL_BUSY
    sat = np.array([[1e4, 1e8, 1e5, 1e5], [1e4, 1e5, 1e8, 1e5], [1e4, 1e5, 1e5, 1e8]], np.float32)
    out, dec = oracle.ann_forward(w, sat)
    assert dec.tolist() == [crn.CH1_OCCUPIED, crn.CH2_OCCUPIED, crn.CH3_OCCUPIED]
    assert np.allclose(np.diag(out), [0.99943, 0.99941, 0.99947], atol=1e-5)


def test_oracle_forward_equals_the_reference_engine(crn, oracle):
    """Golden fixtures: features and Output[1..3] tapped from the UNMODIFIED reference engine (oracle O1)."""
    w = ref_weights(crn)
    files = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        out, dec = oracle.ann_forward(w, z["feat"])
        assert np.array_equal(out, z["ann"]), f          # same libm, same order of operations: bit for bit
        assert np.array_equal(dec, z["decision"]), f
        out2, dec2 = oracle.mlp_f64(crn.config_reference(), z["feat"])
        assert np.abs(out - out2).max() <= 1e-12


def test_oracle_gradient_is_the_gradient(crn, oracle):
    """Central finite differences of E against the back-propagated direction, every one of the 43 weights."""
    rng = np.random.default_rng(5)
    feat, labels = labelled_features(rng, 64)
    tc = crn.ann_train_config(max_epochs=0, init_range=0.0, input_scale=SCALE)
    w = crn.AnnWeights()
    for i in range(5):
        for j in range(1, 6):
            w.wih[i][j] = rng.uniform(-0.5, 0.5)
    for j in range(6):
        for k in range(1, 4):
            w.who[j][k] = rng.uniform(-0.5, 0.5)
    E0, g = oracle.ann_error(w, SCALE, feat, labels)
    assert E0 > 0
    h = 1e-6

    def fd(arr, a, b):
        old = arr[a][b]
        arr[a][b] = old + h
        ep, _ = oracle.ann_error(w, SCALE, feat, labels)
        arr[a][b] = old - h
        em, _ = oracle.ann_error(w, SCALE, feat, labels)
        arr[a][b] = old
        return (ep - em) / (2 * h)

    for i in range(5):
        for j in range(1, 6):
            assert abs(-fd(w.wih, i, j) - g.wih[i][j]) <= 1e-6 * max(1.0, abs(g.wih[i][j])), (i, j)
    for j in range(6):
        for k in range(1, 4):
            assert abs(-fd(w.who, j, k) - g.who[j][k]) <= 1e-6 * max(1.0, abs(g.who[j][k])), (j, k)


def test_oracle_training_learns_the_labels(crn, oracle):
    rng = np.random.default_rng(7)
    feat, labels = labelled_features(rng, 400)             # "about 400 examples" (README.md:104)
    tc = crn.ann_train_config(max_epochs=1500, check_every=100, eta=0.5, alpha=0.9, input_scale=SCALE)
    w0 = crn.AnnWeights()
    E_start, _ = oracle.ann_error(oracle.ann_train(crn.ann_train_config(max_epochs=0, input_scale=SCALE), feat, labels, w0)[0],
                                  [1, 1, 1, 1], feat, labels)
    w, E, epochs = oracle.ann_train(tc, feat, labels, w0)
    assert epochs == 1500 and E < 0.05 * E_start
    out, dec = oracle.ann_forward(w, feat)                   # raw features: the input scale is folded into wih
    assert (dec == labels).mean() >= 0.99
    # early stop at a check point once the target error is reached
    tc2 = crn.ann_train_config(max_epochs=1500, check_every=100, input_scale=SCALE, target_error=10 * E)
    _, E2, ep2 = oracle.ann_train(tc2, feat, labels, w0)
    assert ep2 % 100 == 0 and ep2 < 1500 and E2 <= 10 * E


# ---------------------------------------------------------------- GPU ---------------------------------------

@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "these tests need the B200"
    t.cuda.set_device(0)
    return t


def gpu_forward(crn, torch, w, feat, threshold=0.8):
    n, stride = feat.shape
    d_feat = torch.from_numpy(np.ascontiguousarray(feat, np.float32)).cuda()
    d_out = torch.full((n, 3), float("nan"), dtype=torch.float64, device="cuda")
    d_dec = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    crn.ann_forward(w, d_feat, n, stride, d_out, d_dec, threshold, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_dec.cpu().numpy()


@pytest.mark.gpu
def test_forward_matches_reference_engine_outputs(crn, oracle, torch):
    w = ref_weights(crn)
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz"))):
        z = np.load(f)
        out, dec = gpu_forward(crn, torch, w, z["feat"])
        assert np.abs(out - z["ann"]).max() <= 1e-12, f
        assert np.array_equal(dec, z["decision"]), f


@pytest.mark.gpu
def test_forward_matches_oracle_on_random_features(crn, oracle, torch):
    """100k rows over 14 decades (saturated and mid-sigmoid hidden units alike), wide rows (stride 7), and
    the weights of a freshly trained network."""
    rng = np.random.default_rng(11)
    w = ref_weights(crn)
    n = 100_000
    feat = np.zeros((n, 7), np.float32)
    feat[:, :4] = (10.0 ** rng.uniform(-5, 9, (n, 4))).astype(np.float32)
    feat[: n // 4, :4] = rng.uniform(0, 5, (n // 4, 4)).astype(np.float32)   # unsaturated region
    feat[:, 4:] = np.nan                                                        # columns beyond 4 are not inputs
    out, dec = gpu_forward(crn, torch, w, feat)
    oout, odec = oracle.ann_forward(w, feat)
    assert np.abs(out - oout).max() <= 1e-12
    near = (np.abs(oout - 0.8) <= 1e-9).any(axis=1)
    assert np.array_equal(dec[~near], odec[~near])
    assert len(set(odec.tolist())) == 4                      # every branch of the chain is exercised
    out1, dec1 = gpu_forward(crn, torch, w, feat[:1])         # n = 1, and outputs only / decisions only
    assert np.array_equal(out1, out[:1]) and dec1[0] == dec[0]
    d_feat = torch.from_numpy(feat).cuda()
    d_dec = torch.zeros(n, dtype=torch.int32, device="cuda")
    crn.ann_forward(w, d_feat, n, 7, None, d_dec)
    torch.cuda.synchronize()
    assert np.array_equal(d_dec.cpu().numpy(), dec)


@pytest.mark.gpu
def test_forward_equals_the_fused_kernels_epilogue(crn, oracle, torch):
    """The stand-alone forward pass on the features the fused sensing kernel wrote = the outputs it wrote."""
    cfg = crn.config_welch(1024, 64)
    ng = 64
    iq, _ = oracle.synth(crn.synth_config(cfg.group_samples, dwell_groups=4, snr_db=10.0), ng * cfg.group_samples)
    with crn.Sensor(cfg, device=0) as s:
        d_iq = torch.from_numpy(iq.view(np.float32)).cuda()
        d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
        d_ann = torch.empty(ng, 3, dtype=torch.float64, device="cuda")
        d_dec = torch.empty(ng, dtype=torch.int32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, None, stream)
        d_out2 = torch.empty(ng, 3, dtype=torch.float64, device="cuda")
        d_dec2 = torch.empty(ng, dtype=torch.int32, device="cuda")
        crn.ann_forward(crn.AnnWeights.from_config(cfg), d_feat, ng, cfg.nbands, d_out2, d_dec2, cfg.ann_threshold, 0, stream)
        torch.cuda.synchronize()
    assert torch.equal(d_ann, d_out2) and torch.equal(d_dec, d_dec2)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 37, 400, 100_003])
def test_training_matches_the_cpu_statement(crn, oracle, torch, n):
    rng = np.random.default_rng(100 + n)
    feat, labels = labelled_features(rng, n)
    tc = crn.ann_train_config(max_epochs=230, check_every=50, eta=0.5, alpha=0.9, input_scale=SCALE, seed=n)
    d_feat, d_lab = torch.from_numpy(feat).cuda(), torch.from_numpy(labels).cuda()
    w, E, ep = crn.ann_train(tc, d_feat, 4, d_lab, n)
    ow, oE, oep = oracle.ann_train(tc, feat, labels, crn.AnnWeights())
    assert ep == oep == 230                                  # 4 graph replays of 50 + 30 single epochs
    assert abs(E - oE) <= 1e-11 * abs(oE)
    for a, b in zip(w.arrays(), ow.arrays()):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    # deterministic: the same call again gives the same bits
    w2, E2, _ = crn.ann_train(tc, d_feat, 4, d_lab, n)
    assert E2 == E and all(np.array_equal(a, b) for a, b in zip(w.arrays(), w2.arrays()))


@pytest.mark.gpu
def test_training_continues_from_given_weights_and_stops_at_target(crn, oracle, torch):
    rng = np.random.default_rng(3)
    feat, labels = labelled_features(rng, 400)
    d_feat, d_lab = torch.from_numpy(feat).cuda(), torch.from_numpy(labels).cuda()
    first = crn.ann_train_config(max_epochs=300, check_every=100, input_scale=SCALE)
    w1, E1, _ = crn.ann_train(first, d_feat, 4, d_lab, 400)
    cont = crn.ann_train_config(max_epochs=300, check_every=100, input_scale=SCALE, init_range=0.0)
    w2, E2, ep2 = crn.ann_train(cont, d_feat, 4, d_lab, 400, weights=w1)
    ow2, oE2, _ = oracle.ann_train(cont, feat, labels, w1)
    assert E2 < E1 and abs(E2 - oE2) <= 1e-11 * oE2
    for a, b in zip(w2.arrays(), ow2.arrays()):
        assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    stop = crn.ann_train_config(max_epochs=5000, check_every=100, input_scale=SCALE, target_error=E1)
    _, E3, ep3 = crn.ann_train(stop, d_feat, 4, d_lab, 400)
    assert ep3 % 100 == 0 and ep3 <= 300 and E3 <= E1
    assert crn.lib.crn_ann_train_device(C.byref(stop), None, 4, None, 400, C.byref(w1), None, None, 0, None) == crn.ERR_INVALID
    bad = crn.ann_train_config(alpha=1.0)
    with pytest.raises(crn.CrnError):
        crn.ann_train(bad, d_feat, 4, d_lab, 400)


@pytest.mark.gpu
def test_retrained_predictor_drives_the_fused_kernel(crn, oracle, torch):
    """SURVEY 8f-4 end to end on the GPU: synthetic PU capture -> fused sensing features -> labels from the PU
    state -> retrain -> the fused kernel with the new weights names the occupied channel."""
    cfg = crn.config_welch(1024, 64)
    ng = 600
    sc = crn.synth_config(cfg.group_samples, dwell_groups=3, snr_db=5.0, seed=21, hop_mode=2)
    stream = torch.cuda.current_stream().cuda_stream
    d_iq = torch.empty(ng * cfg.group_samples, 2, dtype=torch.float32, device="cuda")
    d_state = torch.empty(ng, dtype=torch.int32, device="cuda")
    crn.synth_generate(sc, d_iq, 0, ng * cfg.group_samples, d_state, 0, stream)
    d_feat = torch.empty(ng, cfg.nbands, dtype=torch.float32, device="cuda")
    with crn.Sensor(cfg, device=0) as s:
        s.sense_device(d_iq, ng, d_feat, None, None, None, stream)
    d_lab = (d_state + 1).to(torch.int32)                     # PU on channel c -> "CH(c+1) occupied"
    torch.cuda.synchronize()
    fmax = float(d_feat.max())
    tc = crn.ann_train_config(max_epochs=3000, check_every=200, input_scale=[1.0 / fmax] * 4, target_error=0.5)
    w, E, ep = crn.ann_train(tc, d_feat[:400], cfg.nbands, d_lab[:400], 400, stream=stream)   # train on 400 examples
    assert E <= 0.5
    cfg2 = w.into_config(cfg.copy())
    d_dec = torch.empty(ng, dtype=torch.int32, device="cuda")
    d_ann = torch.empty(ng, 3, dtype=torch.float64, device="cuda")
    with crn.Sensor(cfg2, device=0) as s:
        s.sense_device(d_iq, ng, d_feat, d_ann, d_dec, None, stream)
    torch.cuda.synchronize()
    acc = float((d_dec[400:] == d_lab[400:]).float().mean())  # held-out decisions
    assert acc >= 0.99, acc
    # and the CPU statement of the engine agrees with the GPU on those decisions
    iq_head = d_iq[: 8 * cfg.group_samples].cpu().numpy().view(np.complex64).ravel()
    _, oann, odec, _ = oracle.sense_port(cfg2, iq_head)
    assert np.array_equal(odec, d_dec[:8].cpu().numpy()) and np.abs(oann - d_ann[:8].cpu().numpy()).max() <= 1e-5

"""C++ host side (cognitive-radio-network_b200/host): the reference's plugin boundary re-hosted on a
UHD-free radio, the GPU-backed CE_Predictive_Node selected by the scenario's cfg string, crn_replay."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, feat_close

HOST = os.path.join(ROOT, "cognitive-radio-network_b200", "host")
REPLAY = os.path.join(HOST, "crn_replay")
SCENARIO = os.path.join(GOLDEN, "predictive_model_su.cfg")
REF_SCENARIO = os.path.join(GOLDEN, "reference_predictive_model.cfg")   # the reference's scenarios/predictive_model.cfg, unmodified


@pytest.fixture(scope="module")
def replay(crn):
    if not os.path.exists(REPLAY):
        subprocess.run(["make", "-s", "-C", HOST], check=True)
    return REPLAY


def run(replay, args, **kw):
    return subprocess.run([replay] + args, capture_output=True, text=True, timeout=120, **kw)


def test_engine_source_keeps_the_plugin_contract():
    """Same class name / ctor convention / registration string as the reference engine directory."""
    hpp = open(os.path.join(HOST, "cognitive_engines", "CE_Predictive_Node", "CE_Predictive_Node.hpp")).read()
    cpp = open(os.path.join(HOST, "cognitive_engines", "CE_Predictive_Node", "CE_Predictive_Node.cpp")).read()
    assert "class CE_Predictive_Node : public CognitiveEngine" in hpp
    assert "CE_Predictive_Node(int argc, char **argv, ExtensibleCognitiveRadio *_ECR)" in hpp
    assert "virtual void execute();" in hpp
    base = open(os.path.join(HOST, "include", "cognitive_engine.hpp")).read()
    assert "ExtensibleCognitiveRadio *ECR;" in base and "virtual void execute();" in base
    # the arithmetic is not in the engine any more: no FFT, no exp, only libcrnsense calls
    for sym in ("crn_create", "crn_ring_acquire", "crn_submit", "crn_wait"):
        assert sym in cpp
    assert "fft_execute" not in cpp and "exp(" not in cpp


def test_replay_fails_loudly_without_gpu_or_inputs(replay, tmp_path):
    import torch
    r = run(replay, [])
    assert r.returncode == 2
    bad = tmp_path / "bad.cfg"
    bad.write_text("node2 : { cognitive_engine = ; };")
    iq = tmp_path / "z.c64"
    np.zeros(512 * 10, np.complex64).tofile(iq)
    r = run(replay, ["--scenario", str(bad), "--iq", str(iq)])
    assert r.returncode == 1 and "line 1" in r.stderr
    noce = tmp_path / "noce.cfg"
    noce.write_text("node2 : { node_type = \"cognitive radio\"; };")
    r = run(replay, ["--scenario", str(noce), "--iq", str(iq)])
    assert r.returncode == 1 and "must be specified" in r.stderr
    unk = tmp_path / "unk.cfg"
    unk.write_text("// c\nnode2 : { cognitive_engine = \"CE_Nope\"; ce_timeout_ms = 0; };")
    r = run(replay, ["--scenario", str(unk), "--iq", str(iq)])
    assert r.returncode != 0 and "not registered" in r.stdout
    if not torch.cuda.is_available():
        # scenario parses, the engine is found by its cfg string, and then there is no CPU fallback
        r = run(replay, ["--scenario", SCENARIO, "--iq", str(iq), "--packet-len", "512"])
        assert r.returncode != 0
        assert "crn_create failed" in r.stdout and "no usable CUDA device" in r.stdout


REF_ENGINES = "/root/reference/cognitive_engines"


def _make(out, engine_dirs, skip="", extra_inc="", extra_srcs=""):
    cmd = ["make", "-s", "-C", HOST, "OUT=%s" % out, "ENGINE_DIRS=%s" % engine_dirs]
    if skip:
        cmd.append("SKIP=%s" % skip)
    if extra_inc:
        cmd.append("EXTRA_INC=%s" % extra_inc)
    if extra_srcs:
        cmd.append("EXTRA_SRCS=%s" % extra_srcs)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(str(out), "crn_replay")


@pytest.mark.skipif(not os.path.isdir(REF_ENGINES), reason="reference tree not present (GPU box)")
def test_reference_engines_build_unmodified_and_register(crn, tmp_path):
    """the other engines of the reference (CE_Template, CE_TX_CHANNEL_X,
    CE_Random_Behaviour_PU, CE_PU_MARKOV_Chain_Tx) compile from where they lie, unmodified, against this
    radio's headers, and the registration generator wires them up next to the GPU-backed CE_Predictive_Node
    (ours is found first, so the reference's CPU engine of the same name is skipped)."""
    exe = _make(tmp_path, "%s %s" % (os.path.join(HOST, "cognitive_engines"), REF_ENGINES))
    reg = open(os.path.join(str(tmp_path), "lib", "ce_registry_generated.cpp")).read()
    for name in ("CE_Predictive_Node", "CE_Template", "CE_TX_CHANNEL_X", "CE_Random_Behaviour_PU", "CE_PU_MARKOV_Chain_Tx"):
        assert "CRN_REGISTER_CE(%s)" % name in reg
    assert reg.count("CE_Predictive_Node.hpp") == 1 and REF_ENGINES + "/CE_Predictive_Node" not in reg
    # run one of them: CE_Template only looks at CE_metrics.CE_event (CE_Template.cpp:33-60)
    cfg = tmp_path / "t.cfg"
    cfg.write_text('node1 : { cognitive_engine = "CE_Template"; ce_timeout_ms = 0; ce_args = "-d 1"; rx_freq = 833e6; rx_rate = 13e6; };')
    iq = tmp_path / "z.c64"
    np.zeros(512 * 40, np.complex64).tofile(iq)
    r = run(exe, ["--scenario", str(cfg), "--node", "1", "--iq", str(iq), "--packet-len", "512", "--free-run"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "engine=CE_Template" in r.stdout and "40 packets received" in r.stdout
    # the PU engine the shipped scenario uses starts too (it only retunes every 2 s of wall clock)
    cfg.write_text('node1 : { cognitive_engine = "CE_Random_Behaviour_PU"; ce_timeout_ms = 0; tx_freq = 833e6; };')
    r = run(exe, ["--scenario", str(cfg), "--node", "1", "--iq", str(iq), "--packet-len", "512", "--free-run"])
    assert r.returncode == 0 and "final tx_freq=833000000 Hz" in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF_ENGINES), reason="reference tree not present (GPU box)")
def test_reference_cpu_engine_on_this_radio_reproduces_the_fixtures(crn, tmp_path):
    """The radio runtime itself is checked with the reference's UNMODIFIED CPU CE_Predictive_Node plugged into
    it (liquid's FFT replaced by the oracle's restatement - test infrastructure): the lock-step handoff must
    deliver every packet exactly once, in order, so the engine's printed decisions equal the fixtures."""
    oracle_dir = os.path.join(ROOT, "oracle")
    fft_obj = str(tmp_path / "liquid_fft_restated.o")
    subprocess.run(["gcc", "-O2", "-c", "%s/liquid_fft_restated.c" % oracle_dir, "-o", fft_obj], check=True)
    exe = _make(tmp_path, REF_ENGINES, extra_inc="-I %s/compat" % oracle_dir, extra_srcs=fft_obj)
    g = np.load(os.path.join(GOLDEN, "ref_markov_L363.npz"))
    iq = tmp_path / "cap.c64"
    g["iq"].astype(np.complex64).tofile(iq)
    r = run(exe, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "363", "--ce-args", ""])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    import re
    got = []
    for m in re.finditer(r"Channel_State\[(\d)\]: OCCUPIED|ALL BUSY", r.stdout):
        got.append(int(m.group(1)) if m.group(1) else 0)
    assert got == g["decision"].tolist()
    feats = re.findall(r"NOISE FLOOR\s+(\S+)\s+CH1\s+(\S+)\s+CH2\s+(\S+)\s+CH3\s+(\S+)", r.stdout)
    printed = np.array(feats, dtype=np.float64)
    assert np.allclose(printed, g["feat"], rtol=6e-3)   # printed at %.2e upstream (.cpp:207)
    assert "%d forwarded to the CE" % (10 * len(got)) in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF_ENGINES), reason="reference tree not present (GPU box)")
def test_gpu_engine_compiles_against_the_reference_headers():
    """Drop-in check in the other direction: the GPU-backed engine source, unmodified, against the REFERENCE's own
    include/extensible_cognitive_radio.hpp and include/cognitive_engine.hpp (liquid / UHD type names from
    oracle/compat, as for oracle O1), at the reference's language level (makefile:1, -std=c++11)."""
    ref_inc = os.path.join(os.path.dirname(REF_ENGINES), "include")
    src = os.path.join(HOST, "cognitive_engines", "CE_Predictive_Node", "CE_Predictive_Node.cpp")
    r = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "oracle", "compat"),
                        "-I" + ref_inc, "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # and the scenario fixture below IS the reference's file
    ref_cfg = os.path.join(os.path.dirname(REF_ENGINES), "scenarios", "predictive_model.cfg")
    assert open(ref_cfg, "rb").read() == open(REF_SCENARIO, "rb").read()


@pytest.mark.gpu
def test_unmodified_reference_scenario_runs_on_the_gpu_engine(crn, replay, tmp_path):
    """The reference's scenarios/predictive_model.cfg, byte for byte (tests/golden/reference_predictive_model.cfg, a
    reference-held fixture), node 2: scenario reader -> radio setters -> rx/CE workers -> CE_Predictive_Node ->
    libcrnsense, against the reference engine's results on the ragged-frame capture."""
    g = np.load(os.path.join(GOLDEN, "ref_markov_L363.npz"))
    iq = tmp_path / "cap.c64"
    g["iq"].astype(np.complex64).tofile(iq)
    log = tmp_path / "dec.bin"
    r = run(replay, ["--scenario", REF_SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "363",
                     "--ce-args", "-d 0 -q -o %s" % log])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "engine=CE_Predictive_Node" in r.stdout
    nd = len(g["decision"])
    assert "%d packets received, %d forwarded" % (10 * nd, 10 * nd) in r.stdout
    res = (crn.Result * nd).from_buffer_copy(open(log, "rb").read())
    feat, ann, dec, _ = crn.results_to_arrays(res, 4)
    assert feat_close(feat, g["feat"], 1e-4)
    assert np.abs(ann - g["ann"]).max() <= 1e-5
    assert np.array_equal(dec, g["decision"])


def test_unmodified_reference_scenario_parses(replay, tmp_path):
    """Without a GPU the same file must get as far as crn_create (and fail there loudly): every key of both node blocks
    is accepted by the reader."""
    iq = tmp_path / "z.c64"
    np.zeros(512 * 10, np.complex64).tofile(iq)
    r = run(replay, ["--scenario", REF_SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "512"])
    import torch
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "crn_create" in (r.stdout + r.stderr)
    assert "parse error" not in (r.stdout + r.stderr).lower()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ref_markov_L512", "ref_markov_L363", "ref_tone70"])
def test_replayed_capture_matches_reference_engine_fixtures(crn, replay, tmp_path, case):
    """scenario .cfg -> radio -> rx worker -> CE thread -> CE_Predictive_Node::execute() -> libcrnsense,
    against what the reference's unmodified engine produced on the same frames."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    L = int(g["L"])
    iq = tmp_path / "cap.c64"
    g["iq"].astype(np.complex64).tofile(iq)
    log = tmp_path / "dec.bin"
    r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", str(L),
                     "--ce-args", "-d 0 -q -o %s" % log])
    assert r.returncode == 0, r.stdout + r.stderr
    nd = len(g["decision"])
    assert "%d packets received, %d forwarded" % (10 * nd, 10 * nd) in r.stdout
    raw = open(log, "rb").read()
    assert len(raw) == nd * C.sizeof(crn.Result)
    res = (crn.Result * nd).from_buffer_copy(raw)
    feat, ann, dec, _ = crn.results_to_arrays(res, 4)
    assert feat_close(feat, g["feat"], 1e-4)
    assert np.abs(ann - g["ann"]).max() <= 1e-5
    assert np.array_equal(dec, g["decision"])
    assert [r_.first_frame for r_ in res] == [10 * i for i in range(nd)]
    # the retune the engine performed last (CE_Predictive_Node.cpp:245-261)
    last_tx = [crn.TX_FREQ_FOR_DECISION[int(d)] for d in dec if crn.TX_FREQ_FOR_DECISION[int(d)]]
    if last_tx:
        assert "final tx_freq=%.0f Hz" % last_tx[-1] in r.stdout


def test_weight_file_syntax_round_trips_and_bad_files_are_fatal(crn, replay, tmp_path):
    """`-m <file>` takes the reference's own assignment syntax (CE_Predictive_Node.cpp:78-120)."""
    w = crn.AnnWeights.from_config(crn.config_reference())
    text = w.to_literals()
    assert "WeightIH[0][1]   =        -0.18820799999999999;" in text and text.count("Weight") == 43
    w2 = crn.AnnWeights.from_literals(text)
    assert all(np.array_equal(a, b) for a, b in zip(w.arrays(), w2.arrays()))
    # the parser also reads the block exactly as upstream formats it
    up = "  WeightIH[4][5]   =        0.609384;\n  WeightHO[0][1]   =        -7.033320;\n"
    w3 = crn.AnnWeights.from_literals(up)
    assert w3.wih[4][5] == 0.609384 and w3.who[0][1] == -7.033320
    iq = tmp_path / "z.c64"
    np.zeros(512 * 10, np.complex64).tofile(iq)
    bad = tmp_path / "bad_weights.txt"
    bad.write_text("WeightIH[9][1] = 1.0;\n")
    r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--ce-args", "-d 0 -q -m %s" % bad])
    assert r.returncode != 0 and "bad_weights.txt:1" in r.stdout
    r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--ce-args", "-d 0 -q -m /nonexistent"])
    assert r.returncode != 0 and "cannot open weight file" in r.stdout


@pytest.mark.gpu
def test_engine_follows_a_weight_file(crn, replay, tmp_path):
    """Weights handed to the plugin through ce_args reach the GPU: the reference's own literals written to a file
    reproduce the fixture, and a file that flips the sign of the output layer changes the decisions the way the
    CPU statement of the engine says it should."""
    import oracle as O
    g = np.load(os.path.join(GOLDEN, "ref_markov_L512.npz"))
    iq = tmp_path / "cap.c64"
    g["iq"].astype(np.complex64).tofile(iq)
    nd = len(g["decision"])
    w = crn.AnnWeights.from_config(crn.config_reference())
    flipped = crn.AnnWeights.from_literals(w.to_literals())
    for j in range(6):
        for k in range(1, 4):
            flipped.who[j][k] = -flipped.who[j][k]
    for name, weights in (("same", w), ("flipped", flipped)):
        wf = tmp_path / (name + ".txt")
        wf.write_text("// weights written by the test\n" + weights.to_literals())
        log = tmp_path / (name + ".bin")
        r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "512",
                         "--ce-args", "-d 0 -q -o %s -m %s" % (log, wf)])
        assert r.returncode == 0, r.stdout + r.stderr
        res = (crn.Result * nd).from_buffer_copy(open(log, "rb").read())
        feat, ann, dec, _ = crn.results_to_arrays(res, 4)
        oann, odec = O.ann_forward(weights, g["feat"])
        assert np.abs(ann - oann).max() <= 1e-5 and np.array_equal(dec, odec)
        if name == "same":
            assert np.array_equal(dec, g["decision"])
        else:
            assert not np.array_equal(dec, g["decision"])


def test_registration_generator_scans_like_upstream(replay, tmp_path):
    """host/src/config_cognitive_engines.cpp: CE_* directories holding CE_X.hpp + CE_X.cpp register, extra sources of
    an engine directory ride along (src/config_cognitive_engines.cpp:82-107), anything else is ignored; --skip drops
    an engine and the first directory that provides a name wins."""
    gen = os.path.join(HOST, "config_cognitive_engines")
    assert os.path.exists(gen)                      # built by the Makefile before the radio, like upstream's tool
    a, b = tmp_path / "engines_a", tmp_path / "engines_b"
    for d, names in ((a, ["CE_Alpha", "CE_Beta"]), (b, ["CE_Alpha", "CE_Gamma"])):
        for n in names:
            (d / n).mkdir(parents=True)
            (d / n / (n + ".hpp")).write_text("// %s\n" % n)
            (d / n / (n + ".cpp")).write_text("// %s\n" % n)
    (a / "CE_Alpha" / "helper.c").write_text("\n")
    (a / "CE_Alpha" / "notes.txt").write_text("\n")
    (a / "CE_NoHeader").mkdir()
    (a / "CE_NoHeader" / "CE_NoHeader.cpp").write_text("\n")
    (a / "README").write_text("\n")
    out = tmp_path / "lib"
    r = subprocess.run([gen, "--engines", str(a), str(b), "--out", str(out), "--skip", "CE_Beta"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "registered 2 cognitive engine(s): CE_Alpha, CE_Gamma" in r.stdout
    reg = (out / "ce_registry_generated.cpp").read_text()
    assert reg.count("CRN_REGISTER_CE(") == 2 and "CRN_REGISTER_CE(CE_Alpha)" in reg and "CRN_REGISTER_CE(CE_Gamma)" in reg
    assert str(a / "CE_Alpha" / "CE_Alpha.hpp") in reg and str(b / "CE_Alpha") not in reg
    assert "crn_ce_registry_size() { return 2; }" in reg
    mk = (out / "ce_engines.mk").read_text()
    srcs = [l for l in mk.splitlines() if l.startswith("CE_SRCS")][0].split()[2:]
    assert srcs == [str(a / "CE_Alpha" / "CE_Alpha.cpp"), str(a / "CE_Alpha" / "helper.c"),
                    str(b / "CE_Gamma" / "CE_Gamma.cpp")]
    assert subprocess.run([gen, "--bogus"], capture_output=True).returncode == 2


@pytest.mark.skipif(not os.path.isdir(REF_ENGINES), reason="reference tree not present (GPU box)")
def test_lockstep_does_not_hang_on_engines_that_never_sense(crn, tmp_path):
    """Lock-step is crn_replay's default; CE_Template never calls set_ce_sensing(1).  The receiver must not hold its
    packets forever: after the bounded patience they are dropped while sensing is off, as upstream does
    (src/extensible_cognitive_radio.cpp:1310), and the capture ends."""
    exe = _make(tmp_path, "%s %s" % (os.path.join(HOST, "cognitive_engines"), REF_ENGINES))
    cfg = tmp_path / "t.cfg"
    cfg.write_text('node1 : { cognitive_engine = "CE_Template"; ce_timeout_ms = 0; ce_args = "-d 1"; rx_freq = 833e6; rx_rate = 13e6; };')
    iq = tmp_path / "z.c64"
    np.zeros(512 * 40, np.complex64).tofile(iq)
    r = run(exe, ["--scenario", str(cfg), "--node", "1", "--iq", str(iq), "--packet-len", "512", "--lockstep-patience-ms", "200"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "40 packets received, 0 forwarded" in r.stdout


@pytest.mark.gpu
def test_receiver_writes_straight_into_the_pinned_ring(crn, replay, tmp_path):
    """Direct-to-slot handoff (the change to src/extensible_cognitive_radio.cpp:1310-1324): with the GPU engine every
    forwarded packet is received in place in libcrnsense's pinned ring - no rx_buffer -> ce_usrp_rx_buffer -> slot
    copies - in lock-step and in free-run, and the decisions are the reference engine's."""
    g = np.load(os.path.join(GOLDEN, "ref_markov_L512.npz"))
    iq = tmp_path / "cap.c64"
    g["iq"].astype(np.complex64).tofile(iq)
    nd = len(g["decision"])
    log = tmp_path / "dec.bin"
    r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "512", "--ce-args", "-d 0 -q -o %s" % log])
    assert r.returncode == 0, r.stdout + r.stderr
    # the first packet arrives before the engine has created its handle (the provider is registered in execute())
    import re
    direct = int(re.search(r"(\d+) packets received straight into engine slots", r.stdout).group(1))
    assert 10 * nd - 2 <= direct <= 10 * nd, r.stdout
    res = (crn.Result * nd).from_buffer_copy(open(log, "rb").read())
    feat, ann, dec, _ = crn.results_to_arrays(res, 4)
    assert feat_close(feat, g["feat"], 1e-4) and np.array_equal(dec, g["decision"])
    # free-run: packets may be dropped (upstream semantics) but whatever is forwarded is still sensed in order
    r = run(replay, ["--scenario", SCENARIO, "--node", "2", "--iq", str(iq), "--packet-len", "512", "--free-run",
                     "--repeat-packets", "20000", "--ce-args", "-d 0 -q"])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "20000 packets received" in r.stdout

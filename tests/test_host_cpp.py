"""The C++ host layer (cnn-super-resolution_b200/host): DataPipeline / ConfigBasedDataPipeline /
LayerData / Config / the `cnn` CLI, which keep the reference's API and call only the C-ABI.

* CPU: host-only specs (config reader incl. the reference's ConfigTest fixtures, LayerData,
  Argparse grammar, JSON) and CLI argument handling.
* GPU: the spec runner (the reference's spec classes re-stated against the same DataPipeline
  signatures), a ConfigBasedDataPipeline training chain against the committed reference-kernel
  fixture, and the `cnn` CLI end to end on an image against the oracle.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_npz

HOST = os.path.join(ROOT, "cnn-super-resolution_b200", "host")
SPECS = os.path.join(HOST, "bin", "host_specs")
CNN = os.path.join(HOST, "bin", "cnn")


@pytest.fixture(scope="module", autouse=True)
def built():
    import _pkg
    _pkg.load().build()
    subprocess.check_call(["make", "-C", HOST, "-s"])


def run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)


def test_host_only_specs():
    r = run([SPECS, "hostonly", os.path.join(GOLDEN, "ref_config")])
    assert r.returncode == 0, r.stdout
    assert "PASSED" in r.stdout


def test_cli_argument_errors(tmp_path):
    """reference: src/Main_cl.cpp:43-72 -- missing required option / missing --out without dry"""
    r = run([CNN, "-i", "x.ppm"])
    assert r.returncode != 0 and "config" in r.stdout
    r = run([CNN, "-c", "cfg.json", "-i", "x.ppm"])
    assert r.returncode != 0 and "Either provide out path or do the dry run" in r.stdout
    r = run([CNN, "help"])
    assert r.returncode == 0 and "--config" in r.stdout
    # a config that fails Config::validate (even f1) -> error message, not a crash
    bad = tmp_path / "bad.json"
    bad.write_text('{"n1": 8, "n2": 4, "f1": 8, "f2": 1, "f3": 5, "momentum": 0.9, '
                   '"weight_decay_parameter": 0.0, "learning_rates": [1, 1, 1]}')
    r = run([CNN, "dry", "-c", str(bad), "-i", "x.ppm"])
    assert r.returncode != 0 and "f1 should be odd" in r.stdout
    # unparsable config -> IOException text
    r = run([CNN, "dry", "-c", os.path.join(GOLDEN, "ref_config", "config_non_parseable.json"),
             "-i", "x.ppm"])
    assert r.returncode != 0 and "Json parsing error" in r.stdout


def write_config(path, cfg, params_file=""):
    n1, n2, f1, f2, f3 = cfg
    d = {"n1": n1, "n2": n2, "f1": f1, "f2": f2, "f3": f3, "momentum": 0.9,
         "weight_decay_parameter": 0.001, "learning_rates": [1e-3, 1e-3, 1e-4],
         "parameters_file": params_file}
    for i in (1, 2, 3):
        d["parameters_distribution_%d" % i] = {"mean_w": 0.0, "mean_b": 0.0,
                                               "std_deviation_w": 0.01, "std_deviation_b": 0.0}
    with open(path, "w") as fh:
        json.dump(d, fh)


def write_params(path, params, epochs=0):
    d = {"epochs": epochs}
    for l in (1, 2, 3):
        d["layer%d" % l] = {"weights": [float(np.float32(v)) for v in params["w%d" % l]],
                            "bias": [float(np.float32(v)) for v in params["b%d" % l]]}
    with open(path, "w") as fh:
        json.dump(d, fh)


@pytest.mark.gpu
def test_spec_runner_on_gpu():
    r = run([SPECS, "specs", GOLDEN, os.path.join(GOLDEN, "ref_config")])
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("[+]") == 12, r.stdout   # the reference's 11 specs (ConfigTest = 1) + error behaviour


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ref_train_chain.npz", "ref_train_915.npz"])
def test_config_based_pipeline_training_chain(tmp_path, name):
    """ConfigBasedDataPipeline::execute_batch + update_parameters for 2 epochs x 2 chunks ==
    what the reference's kernels give (per-epoch weights <= 1e-4 relative; the parameter file
    itself has 6 significant digits, quirk Q6), epochs counter == number of updates (Q5)."""
    g = load_npz(name)
    cfg = tuple(int(v) for v in g["cfg"])
    ns, w, h, chunk, epochs = (int(v) for v in g["dims"])
    params = {k[3:]: g[k] for k in g if k.startswith("p0_")}
    pfile, cfile = str(tmp_path / "params.json"), str(tmp_path / "config.json")
    write_params(pfile, params, epochs=7)
    write_config(cfile, cfg, pfile)
    inp = {"config_path": cfile, "w": w, "h": h, "epochs": epochs, "chunk": chunk,
           "x": [float(v) for v in g["x"].reshape(-1)], "gt": [float(v) for v in g["gt"].reshape(-1)]}
    with open(tmp_path / "in.json", "w") as fh:
        json.dump(inp, fh)
    out = str(tmp_path / "out.json")
    r = run([SPECS, "chain", str(tmp_path / "in.json"), out])
    assert r.returncode == 0, r.stdout
    res = json.load(open(out))
    assert res["epochs"] == 7 + epochs
    for l in (1, 2, 3):
        np.testing.assert_allclose(np.array(res["layer%d" % l]["weights"], np.float32),
                                   g["e%d_w%d" % (epochs, l)], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(np.array(res["layer%d" % l]["bias"], np.float32),
                                   g["e%d_b%d" % (epochs, l)], rtol=1e-4, atol=1e-7)
    sse = float(r.stdout.split("validation_sse")[1].split()[0])
    assert sse == pytest.approx(float(g["final_sse"]), rel=1e-4)


def write_ppm(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "wb") as fh:
        fh.write(b"P6\n%d %d\n255\n" % (w, h))
        fh.write(np.ascontiguousarray(rgb, np.uint8).tobytes())


def read_ppm(path):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6"
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


@pytest.mark.gpu
def test_cnn_cli_forward_image(tmp_path, port):
    """`cnn -c cfg -i in.ppm -o out.ppm` == oracle pipeline: extract_luma(/255) ->
    subtract_mean (quirk Q1: mean of squares) -> 3 layers -> swap_luma."""
    from helpers import make_params
    from oracle.loader import NetState
    rng = np.random.default_rng(77)
    cfg = (64, 32, 9, 1, 5)
    h, w = 70, 93
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([127 + 100 * np.sin(xx / 9.0), 127 + 100 * np.cos(yy / 7.0),
                    127 + 60 * np.sin((xx + yy) / 5.0)], -1) + rng.normal(0, 8, (h, w, 3))
    rgb = np.clip(rgb, 0, 255).astype(np.uint8)
    params = make_params(rng, *cfg)
    pfile, cfile = str(tmp_path / "params.json"), str(tmp_path / "config.json")
    write_params(pfile, params)
    write_config(cfile, cfg, pfile)
    write_ppm(tmp_path / "in.ppm", rgb)
    r = run([CNN, "-c", cfile, "-i", str(tmp_path / "in.ppm"), "-o", str(tmp_path / "out.ppm")])
    assert r.returncode == 0, r.stdout
    got = read_ppm(tmp_path / "out.ppm")
    # oracle
    rgba = np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], -1)
    luma = port.extract_luma(rgba, True)
    port.subtract_mean(luma.reshape(-1), with_event=True)
    p32 = {k: np.array([np.float32(float(np.float32(v))) for v in a], np.float32)
           for k, a in params.items()}
    net = NetState(*cfg, p32)
    _, _, o3 = port.net_forward(net, luma, w, h, 1)
    exp = port.swap_luma(rgba, o3[0], w - 12, h - 12)
    diff = np.abs(got.astype(int) - exp.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.gpu
def test_cnn_cli_training_and_profile_output(tmp_path):
    """`cnn train ... profile`: 80/20 split, one update per epoch, parameters file written with
    the epoch counter, validation line printed, per-kernel totals in the reference's line format
    (src/opencl/Context.cpp:93-95, parsed by the reference's profile.py:9)."""
    import re
    rng = np.random.default_rng(5)
    d = tmp_path / "samples"
    d.mkdir()
    for i in range(10):
        gt = rng.integers(0, 256, (33, 33, 1)).astype(np.uint8).repeat(3, -1)
        write_ppm(d / ("sample_%d_large.ppm" % i), gt)
        write_ppm(d / ("sample_%d_small.ppm" % i), np.clip(gt.astype(int) + rng.integers(-9, 9, gt.shape), 0, 255))
    (d / "notes.txt").write_text("not an image")
    cfile = str(tmp_path / "config.json")
    write_config(cfile, (8, 4, 9, 1, 5))
    out = str(tmp_path / "trained.json")
    env = dict(os.environ, CNN_SR_SEED="3")
    r = run([CNN, "train", "profile", "-c", cfile, "-i", str(d), "-o", out, "--epochs", "3"], env=env)
    assert r.returncode == 0, r.stdout
    assert "validation_set_size: 2/10 = 20%" in r.stdout
    assert "mini-batch size: 6" in r.stdout            # 8/2 + 2  (src/Main_cl.cpp:128-129)
    assert "mean validation error" in r.stdout and "DONE" in r.stdout
    assert "is not a sample image" in r.stdout
    res = json.load(open(out))
    assert res["epochs"] == 3 and len(res["layer1"]["weights"]) == 81 * 8
    prof = re.findall(r"Kernel '.*/(.*?]).*?([\-e.\d]+)ns.*?([\-e.\d]+)s", r.stdout)  # profile.py:9
    names = " ".join(p[0] for p in prof)
    for k in ("layer_uber_kernel.cl", "layer_deltas.cl", "backpropagate.cl", "update_parameters.cl"):
        assert k in names, r.stdout
    assert sum(int(float(p[1])) for p in prof) > 0


def _write_samples(d, rng, n):
    d.mkdir()
    for i in range(n):
        gt = rng.integers(0, 256, (33, 33, 1)).astype(np.uint8).repeat(3, -1)
        write_ppm(d / ("sample_%d_large.ppm" % i), gt)
        write_ppm(d / ("sample_%d_small.ppm" % i),
                  np.clip(gt.astype(int) + rng.integers(-9, 9, gt.shape), 0, 255))


@pytest.mark.gpu
def test_cnn_cli_resume_state(tmp_path):
    """the parameters file carries an optional "resume" key (9-digit parameters +
    the momentum state, src/ConfigBasedDataPipeline.cpp:432-465 writes neither).  Training 2
    epochs in one run == 1 epoch, save, resume, 1 more epoch; without the key (the reference's
    format) momentum restarts at zero and 6-digit weights differ."""
    rng = np.random.default_rng(11)
    d = tmp_path / "samples"
    _write_samples(d, rng, 4)              # 4 samples: the 20 % validation set is empty
    cfg = (8, 4, 9, 1, 5)
    c0 = str(tmp_path / "c0.json")
    write_config(c0, cfg)
    env = dict(os.environ, CNN_SR_SEED="3")

    def train(config, out, epochs, extra_env=None):
        e = dict(env, **(extra_env or {}))
        r = run([CNN, "train", "-c", config, "-i", str(d), "-o", out, "--epochs", str(epochs)], env=e)
        assert r.returncode == 0, r.stdout
        return json.load(open(out)), r.stdout

    init, _ = train(c0, str(tmp_path / "init.json"), 0)       # the seeded random initialisation
    assert init["epochs"] == 0 and "resume" in init
    c1 = str(tmp_path / "c1.json")
    write_config(c1, cfg, str(tmp_path / "init.json"))
    two, _ = train(c1, str(tmp_path / "two.json"), 2)
    one, _ = train(c1, str(tmp_path / "one.json"), 1)
    assert any(abs(v) > 0 for v in one["resume"]["layer1"]["previous_delta_w"])
    c2 = str(tmp_path / "c2.json")
    write_config(c2, cfg, str(tmp_path / "one.json"))
    resumed, log = train(c2, str(tmp_path / "resumed.json"), 1)
    assert "Resume state found" in log and resumed["epochs"] == 2
    for l in ("layer1", "layer2", "layer3"):
        for k in ("weights", "bias", "previous_delta_w", "previous_delta_b"):
            np.testing.assert_allclose(np.array(resumed["resume"][l][k]), np.array(two["resume"][l][k]),
                                       rtol=2e-6, atol=1e-9, err_msg="%s %s" % (l, k))
    # the reference's own format: no resume key -> momentum restarts, result differs
    plain, _ = train(c1, str(tmp_path / "plain.json"), 1, {"CNN_SR_RESUME_STATE": "0"})
    assert "resume" not in plain and set(plain) == {"epochs", "layer1", "layer2", "layer3"}
    c3 = str(tmp_path / "c3.json")
    write_config(c3, cfg, str(tmp_path / "plain.json"))
    cold, log = train(c3, str(tmp_path / "cold.json"), 1)
    assert "Resume state found" not in log
    a = np.array(cold["resume"]["layer1"]["weights"])
    b = np.array(two["resume"]["layer1"]["weights"])
    assert np.abs(a - b).max() > 1e-7


@pytest.mark.gpu
def test_cnn_cli_data_parallel_two_gpus(tmp_path):
    """`cnn train` on 2 GPUs (CNN_SR_WORLD=2, one process per GPU, NCCL id through
    CNN_SR_COMM_FILE) must write the same parameters as the 1-GPU run of the same samples."""
    import ctypes
    try:
        n = ctypes.c_int(0)
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n))
    except OSError:
        pytest.skip("no CUDA runtime")
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(12)
    d = tmp_path / "samples"
    _write_samples(d, rng, 12)
    cfg = (8, 4, 9, 1, 5)
    c0 = str(tmp_path / "c0.json")
    write_config(c0, cfg)
    env = dict(os.environ, CNN_SR_SEED="5")
    r = run([CNN, "train", "-c", c0, "-i", str(d), "-o", str(tmp_path / "init.json"), "--epochs", "0"], env=env)
    assert r.returncode == 0, r.stdout
    c1 = str(tmp_path / "c1.json")
    write_config(c1, cfg, str(tmp_path / "init.json"))
    r = run([CNN, "train", "-c", c1, "-i", str(d), "-o", str(tmp_path / "single.json"), "--epochs", "3"], env=env)
    assert r.returncode == 0, r.stdout
    procs = []
    for rank in range(2):
        e = dict(env, CNN_SR_WORLD="2", CNN_SR_RANK=str(rank), CNN_SR_COMM_FILE=str(tmp_path / "nccl.id"))
        procs.append(subprocess.Popen([CNN, "train", "-c", c1, "-i", str(d), "-o", str(tmp_path / "dp.json"),
                                       "--epochs", "3"], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DATA PARALLEL: rank 0 of 2" in outs[0] and "mean validation error" in outs[0]
    single, dp = json.load(open(tmp_path / "single.json")), json.load(open(tmp_path / "dp.json"))
    assert dp["epochs"] == 3
    for l in ("layer1", "layer2", "layer3"):
        for k in ("weights", "bias"):
            np.testing.assert_allclose(np.array(dp["resume"][l][k]), np.array(single["resume"][l][k]),
                                       rtol=2e-5, atol=1e-9, err_msg="%s %s" % (l, k))
    v1 = [l for l in r.stdout.splitlines() if "mean validation error" in l]
    v2 = [l for l in outs[0].splitlines() if "mean validation error" in l]
    assert len(v1) == len(v2) > 0
    for a, b in zip(v1, v2):
        fa, fb = float(a.split("error:")[1].split()[0]), float(b.split("error:")[1].split()[0])
        assert fa == pytest.approx(fb, rel=1e-4)

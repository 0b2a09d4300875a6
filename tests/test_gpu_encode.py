"""GPU tests of the caller side of the path kept on the device: tdb200_encode_batch
(TurboEnCoding, ITTC/log_map.cpp:700-730) bit-exact against the oracle's restatement and, where the
reference build travelled (oracle/_ref), against the reference itself; tdb200_channel_batch
(module + AWGN + demodule, ITTC/main.cpp:197-202) by its statistics."""
import numpy as np
import pytest

from oracle_lib import RefLib

pytestmark = pytest.mark.gpu


def _torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


@pytest.mark.parametrize("K", [40, 48, 104, 512, 1008, 2048, 4160, 6144])
def test_encoder_bit_exact(oracle, K):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    rng = np.random.default_rng(K)
    n_cb = 5
    bits = rng.integers(0, 2, size=(n_cb, K), dtype=np.uint8)
    bits[0] = 0
    bits[0, 0] = 1          # impulse: the reference's known answer 111011011011... / tail 000111000111 for K = 6144
    bits[1] = 1
    pi = oracle.qpp(K)
    dec = TurboDecoder(K, algo="maxlog_s16", max_batch=8)
    dev_out = dec.encode(torch.from_numpy(bits).cuda()).cpu().numpy()
    host_out = dec.encode(bits)
    ref = RefLib(K, *oracle.lte_params(K)) if RefLib.available() else None
    for c in range(n_cb):
        want = oracle.encode(bits[c].astype(np.int32), pi).astype(np.uint8)
        assert np.array_equal(dev_out[c], want), "device path, cb %d" % c
        assert np.array_equal(host_out[c], want), "host path, cb %d" % c
        if ref is not None:
            assert np.array_equal(want, ref.encode(bits[c].astype(np.int32)).astype(np.uint8))
    if K == 6144:
        assert "".join(map(str, dev_out[0][:12])) == "111011011011"


def test_channel_statistics_and_determinism(oracle):
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 6144, 64
    dec = TurboDecoder(K, algo="maxlog_s16", max_batch=64)
    sigma = oracle.sigma(1.0, K)
    coded = torch.randint(0, 2, (n_cb, 3 * K + 12), dtype=torch.uint8, device="cuda")
    a = dec.channel(coded, sigma, seed=11)
    b = dec.channel(coded, sigma, seed=11)
    c = dec.channel(coded, sigma, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # LLR = 2 (x + sigma n) / sigma^2  ->  n recovered from the known x is standard normal
    x = coded.float() * 2 - 1
    n = (a * (sigma * sigma / 2) - x) / sigma
    N = n.numel()
    assert abs(float(n.mean())) < 5 / N ** 0.5
    assert abs(float(n.var()) - 1.0) < 5 * (2.0 / N) ** 0.5
    assert abs(float((n ** 4).mean()) - 3.0) < 0.05          # kurtosis of a Gaussian (a 12-term CLT sum gives 2.9)
    assert float(n.abs().max()) > 4.5                        # real tails: a CLT sum of 12 uniforms cannot exceed 6
    # host path and float64 output agree with the device path
    h = dec.channel(coded[:3].cpu().numpy(), sigma, seed=11)
    assert np.array_equal(h, a[:3].cpu().numpy())
    d64 = dec.channel(coded[:3], sigma, seed=11, dtype="float64")
    assert np.allclose(d64.cpu().numpy(), a[:3].cpu().numpy(), rtol=1e-6)
    d16 = dec.channel(coded[:3], sigma, seed=11, dtype="float16")
    assert torch.equal(d16, a[:3].half())


def test_encode_channel_decode_round_trip(oracle):
    """Size-independent property at the BASELINE size: random bits -> device encoder -> device channel at
    1.5 dB -> decoder returns the bits (every mode)."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder
    K, n_cb = 6144, 32
    bits = torch.randint(0, 2, (n_cb, K), dtype=torch.uint8, device="cuda")
    for algo, dt in (("maxlog_s16", "float32"), ("maxlog_s16", "float16"), ("maxlog_f32", "float32"), ("linlogmap_f32", "float16"),
                     ("logmap_f32", "float32"), ("logmap_f64", "float64"), ("logmap_f64", "float16")):
        dec = TurboDecoder(K, n_iter=8, algo=algo, max_batch=32)
        llr = dec.channel(dec.encode(bits), oracle.sigma(1.5, K), seed=5, dtype=dt)
        out = dec.decode(llr, want=("bits",))
        assert torch.equal(out["bits"], bits), algo


def test_cpp_multi_gpu_caller():
    """compat/tdb200_burst.cpp: a C++ program over the C ABI (no Python, no torch in that process), one host
    thread per visible GPU, codeblock-sharded.  At 2 dB nothing may be in error."""
    _torch_cuda()
    import json
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "turbo_decoder_cuda_b200", "lib", "tdb200_burst")
    if not os.path.exists(exe):
        pytest.skip("tdb200_burst not built")
    out = subprocess.run([exe, "--total", "3001", "--chunk", "1024", "--ebn0", "2.0", "--early-term", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["codeblocks"] == 3001 and r["bit_errors"] == 0 and r["frame_errors"] == 0
    assert 2.0 <= r["mean_iters"] < 8.0 and r["gbit_s"] > 1.0
    # main.cpp's loop with the rate-matching and mapping stages restored (16QAM, rate 1/2), all through the C ABI
    out = subprocess.run([exe, "--total", "1500", "--chunk", "1024", "--ebn0", "5.0", "--modulation", "4", "--E", "12288", "--early-term", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["codeblocks"] == 1500 and r["modulation"] == 4 and r["E"] == 12288 and r["bit_errors"] == 0
    assert subprocess.run([exe, "--modulation", "5"], capture_output=True).returncode == 2
    # the CRC stopping rule from C++: blocks get a CRC24B attached on the device, fewer iterations than rule 1
    its = {}
    for rule in ("1", "2"):
        out = subprocess.run([exe, "--total", "2048", "--chunk", "1024", "--ebn0", "1.5", "--early-term", rule],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        r = json.loads(out.stdout.strip().splitlines()[-1])
        assert r["bit_errors"] == 0
        its[rule] = r["mean_iters"]
    assert its["2"] < its["1"]


def test_device_path_is_cuda_graph_capturable():
    """With device buffers the three entry points only enqueue kernels on the caller's stream, so a whole
    encode -> channel -> decode step can be captured once and replayed as a CUDA graph."""
    torch = _torch_cuda()
    from turbo_decoder_cuda_b200 import TurboDecoder, decoder as tdb
    from turbo_decoder_cuda_b200.synth import sigma_from_ebn0
    K, n = 6144, 64
    dec = TurboDecoder(K, n_iter=8, algo="maxlog_s16", max_batch=n)
    bits = torch.randint(0, 2, (n, K), dtype=torch.uint8, device="cuda")
    coded = torch.empty((n, 3 * K + 12), dtype=torch.uint8, device="cuda")
    llr = torch.empty((n, 3 * K + 12), dtype=torch.float32, device="cuda")
    out = torch.zeros((n, K), dtype=torch.uint8, device="cuda")
    L = dec._L
    sigma = sigma_from_ebn0(1.5, K)

    def step(stream):
        assert L.tdb200_encode_batch(dec._h, bits.data_ptr(), coded.data_ptr(), tdb.MEM_DEVICE, n, stream) == 0
        assert L.tdb200_channel_batch(dec._h, coded.data_ptr(), llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, n, sigma, 3, stream) == 0
        dec.decode_raw(llr.data_ptr(), tdb.LLR_F32, tdb.MEM_DEVICE, n, bits=out.data_ptr(), stream=stream)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step(s.cuda_stream)          # warm-up outside the capture (lazy module loading)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        bits.copy_(torch.randint(0, 2, (n, K), dtype=torch.uint8, device="cuda"))
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, bits)

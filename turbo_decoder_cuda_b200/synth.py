"""Synthetic traffic for the decoder: batched PCCC encoder + BPSK/AWGN channel + soft demapper.

This is the harness side of the hot path: it restates, vectorised over the
batch dimension with torch ops (CPU or CUDA tensors), what the reference does per frame in
    rsc_encode / encoderm_turbo   ITTC/log_map.cpp:451-583   (encoder, tail, multiplex order)
    module / AWGN / demodule      ITTC/main.cpp:197-202, ITTC/modanddem.cpp:189-224
so that bench.py and the BER harness can make LLR batches without touching oracle/.
It is not part of the decode path and never runs inside a timed region.
"""
import math

import torch

from .decoder import lte_qpp_params


def qpp_permutation(K, f1=None, f2=None, device="cpu"):
    """pi(i) = (f1*i + f2*i^2) mod K  (gen_qpp_index, ITTC/log_map.cpp:616-624)."""
    if f1 is None:
        f1, f2 = lte_qpp_params(K)
    i = torch.arange(K, dtype=torch.int64)
    return ((f1 * i + ((f2 * i) % K) * i) % K).to(device)


def _rsc_encode(bits):
    """(13,15)_8 RSC with trellis termination over a [B, K] uint8 tensor.

    Returns (parity[B, K+3], tail_systematic[B, 3]).  Registers (s0 newest): feedback taps 1011,
    forward taps 1101 (ITTC/log_map.h:34-36); tail input = s1 ^ s2 so the feedback sum is zero
    (log_map.cpp:483-491).
    """
    B, K = bits.shape
    dev = bits.device
    s0 = torch.zeros(B, dtype=torch.uint8, device=dev)
    s1 = torch.zeros_like(s0)
    s2 = torch.zeros_like(s0)
    par = torch.empty((B, K + 3), dtype=torch.uint8, device=dev)
    tail = torch.empty((B, 3), dtype=torch.uint8, device=dev)
    for i in range(K + 3):
        if i < K:
            d = bits[:, i]
        else:
            d = s1 ^ s2
            tail[:, i - K] = d
        a = d ^ s1 ^ s2
        par[:, i] = a ^ s0 ^ s2
        s2, s1, s0 = s1, s0, a
    return par, tail


def turbo_encode(bits, pi):
    """bits [B, K] (0/1) -> coded [B, 3K+12] uint8 in the reference's multiplex order
    (log_map.cpp:566-578): [3i]=sys, [3i+1]=par1, [3i+2]=par2, then (x,z)x3 of RSC1, (x',z')x3 of RSC2."""
    bits = bits.to(torch.uint8)
    B, K = bits.shape
    p1, t1 = _rsc_encode(bits)
    p2, t2 = _rsc_encode(bits[:, pi])
    out = torch.empty((B, 3 * K + 12), dtype=torch.uint8, device=bits.device)
    out[:, 0:3 * K:3] = bits
    out[:, 1:3 * K:3] = p1[:, :K]
    out[:, 2:3 * K:3] = p2[:, :K]
    out[:, 3 * K + 0:3 * K + 6:2] = t1
    out[:, 3 * K + 1:3 * K + 6:2] = p1[:, K:]
    out[:, 3 * K + 6:3 * K + 12:2] = t2
    out[:, 3 * K + 7:3 * K + 12:2] = p2[:, K:]
    return out


def sigma_from_ebn0(ebn0_db, K):
    """Noise std for BPSK at the code rate K/(3K+12)  (ITTC/main.cpp:47,174)."""
    rate = K / (3.0 * K + 12.0)
    return 10.0 ** (-ebn0_db / 20.0) * math.sqrt(0.5 / rate)


def channel_llr(coded, sigma, generator=None, dtype=torch.float32):
    """BPSK (+1 for bit 1) + AWGN + LLR = 2 r / sigma^2  (demodule with Kf = 1/(2 sigma^2))."""
    x = coded.to(torch.float32) * 2.0 - 1.0
    noise = torch.randn(x.shape, generator=generator, device=x.device, dtype=torch.float32)
    return ((x + sigma * noise) * (2.0 / (sigma * sigma))).to(dtype)


_handles = {}


def _handle(K, index):
    """A decoder handle used only for its encoder / channel entry points (cached per block size and device)."""
    from .decoder import TurboDecoder
    key = (K, index)
    if key not in _handles:
        _handles[key] = TurboDecoder(K, n_iter=1, algo="maxlog_s16", device=index, max_batch=2)
    return _handles[key]


def make_batch(K, n_cb, ebn0_db, seed=0, device="cpu", dtype=torch.float32, chunk=4096, kernels=True):
    """Seeded random codeblocks through the channel: (bits [n_cb,K] uint8, llr [n_cb,3K+12]).

    On a CUDA device the encoder and the channel are the library's own kernels
    (tdb200_encode_batch / tdb200_channel_batch, csrc/tdb200_encode.cu); on the CPU, or with
    kernels=False, the batched torch restatement above.  The two paths draw different noise."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sigma = sigma_from_ebn0(ebn0_db, K)
    if dev.type == "cuda" and kernels and dtype in (torch.float32, torch.float64):
        h = _handle(K, dev.index if dev.index is not None else torch.cuda.current_device())
        bits = torch.randint(0, 2, (n_cb, K), generator=g, device=dev, dtype=torch.uint8)
        llr = h.channel(h.encode(bits), sigma, seed=seed, dtype=str(dtype).replace("torch.", ""))
        return bits, llr
    pi = qpp_permutation(K, device=dev)
    bits_all, llr_all = [], []
    for c0 in range(0, n_cb, chunk):
        n = min(chunk, n_cb - c0)
        bits = torch.randint(0, 2, (n, K), generator=g, device=dev, dtype=torch.uint8)
        llr_all.append(channel_llr(turbo_encode(bits, pi), sigma, g, dtype))
        bits_all.append(bits)
    return torch.cat(bits_all), torch.cat(llr_all)

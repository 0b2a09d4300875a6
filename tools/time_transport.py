"""Device time of the CRC kernels and transport-block-level decode throughput (CUDA events).

    python tools/time_transport.py [--json out.json]

4096 code blocks of K = 6144: CRC24B check / attach; then 315 transport blocks of the largest LTE size
(A = 75376 -> 13 code blocks of 5824) through segment -> encode -> channel -> decode (+ CRC24B per code
block, CRC24A per transport block, concatenation), Eb/N0 1.5 dB, early termination on.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbo_decoder_cuda_b200 import TurboDecoder  # noqa: E402
from turbo_decoder_cuda_b200.decoder import CRC24B  # noqa: E402
from turbo_decoder_cuda_b200.synth import sigma_from_ebn0  # noqa: E402
from turbo_decoder_cuda_b200.transport import TransportBlockCodec  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    a = ap.parse_args()
    K, N = 6144, 4096
    dec = TurboDecoder(K, max_batch=N)
    bits = torch.randint(0, 2, (N, K), dtype=torch.uint8, device="cuda")
    res = {"crc24b_attach_ms": timed(lambda: dec.crc24_attach(bits, CRC24B)),
           "crc24b_check_ms": timed(lambda: dec.crc24_check(bits, CRC24B)), "codeblocks": N, "K": K}
    res["crc24b_check_gb_s"] = N * K / res["crc24b_check_ms"] / 1e6
    A, n_tb = 75376, 315
    tb = TransportBlockCodec(A, n_iter=8, early_term=True, max_batch=n_tb * 13)
    payload = torch.randint(0, 2, (n_tb, A), dtype=torch.uint8, device="cuda")
    blocks = tb.segment(payload)
    coded = tb.encode(blocks)
    Kp = tb.seg["K_plus"]
    sigma = sigma_from_ebn0(1.5, Kp)
    llrs = [(k, tb.dec[k].channel(cw, sigma, seed=1)) for k, cw in coded]
    out, tb_ok, cb_ok = tb.decode(llrs)
    ms = timed(lambda: tb.decode(llrs), reps=5)
    ms_dec = timed(lambda: [tb.dec[k].decode(l) for k, l in llrs], reps=5)
    res.update({"transport_block_bits": A, "transport_blocks": n_tb, "code_blocks": n_tb * tb.seg["C"], "K_plus": Kp,
                "ebn0_db": 1.5, "tb_decode_ms": ms, "decode_only_ms": ms_dec,
                "tb_payload_gbit_s": n_tb * A / ms / 1e6, "tb_ok_fraction": float(tb_ok.float().mean()),
                "payload_bit_errors": int((out != payload).sum())})
    print(json.dumps(res))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

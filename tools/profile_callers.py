"""One launch of every caller-side kernel at the BASELINE block size, for an ncu capture:
    ncu --set full --clock-control none -k regex:"demap|rate_|crc24|modulate|awgn" -o gpurun_out/prof_callers python tools/profile_callers.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from turbo_decoder_cuda_b200 import TurboDecoder  # noqa: E402
from turbo_decoder_cuda_b200.decoder import CRC24B  # noqa: E402

K, N = 6144, 2048
dec = TurboDecoder(K, n_iter=8, max_batch=N)
bits = torch.randint(0, 2, (N, K), dtype=torch.uint8, device="cuda")
dec.crc24_attach(bits, CRC24B)                                   # crc24_kernel (attach)
coded = dec.encode(bits)
si, sq = dec.modulate(coded, 6)                                  # modulate_kernel<6>
sigma = 0.14
ri, rq = dec.awgn(si, sigma, seed=1), dec.awgn(sq, sigma, seed=2)  # awgn_kernel
kf = 1.0 / (2 * sigma * sigma)
out = dec.decode_symbols(ri, rq, 6, kf)["bits"]                  # demap32_kernel<6, float, S8> + fast_s16_kernel
print("64QAM bit errors", int((out != bits).sum()), "crc ok", int(dec.crc24_check(out, CRC24B).sum()), "of", N)  # crc24_kernel (check)
l64 = dec.demap(ri[:256].double(), rq[:256].double(), 6, kf, dtype="float64")      # demap64_kernel<6>
tx = dec.rate_match(coded, 2 * K, 0)                             # rate_match_kernel (rate 1/2)
e_llr = (tx.float() * 2 - 1) * 4.0
out = dec.decode_rm(e_llr, 0)["bits"]                            # rate_dematch_kernel<F32, S8> + fast_s16_kernel
print("rate-1/2 bit errors", int((out != bits).sum()))
llr = dec.rate_dematch(e_llr, 0)                                 # rate_dematch_kernel<F32, F32>
torch.cuda.synchronize()

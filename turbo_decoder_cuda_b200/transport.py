"""Transport-block caller of the decode path: CRC24A attachment, code-block
segmentation with filler bits and a CRC24B per block, and the inverse after decoding -- TS 36.212
5.1.1 / 5.1.2.  The reference has only a placeholder for this stage (previous/Decoder.cc:1026,
:1098-1099).  Everything that computes on bits runs on the device through the C ABI
(tdb200_crc24_*, tdb200_encode_batch, tdb200_decode_batch); this module only slices tensors.

    tb = TransportBlockCodec(A)                     # A payload bits per transport block
    blocks = tb.segment(payload)                    # [(K, uint8 [n_tb * C_K, K]), ...] with CRCs attached
    coded = tb.encode(blocks)                       # [(K, uint8 [n, 3K+12]), ...]
    ...channel...
    payload_hat, tb_ok, cb_ok = tb.decode(llrs)     # llrs: [(K, [n, 3K+12]), ...]
"""
import torch

from .decoder import CRC24A, CRC24B, TurboDecoder, segmentation


class TransportBlockCodec:
    def __init__(self, A, device=0, n_iter=8, algo="maxlog_s16", early_term=True, max_batch=0):
        """early_term: True = the CRC stopping rule (the packed 16-bit decoders), "hda" = decisions + magnitude, False = none."""
        self.A = int(A)
        self.B = self.A + 24                                    # with the transport-block CRC24A
        s = self.seg = segmentation(self.B)
        self.device = torch.device("cuda", device)
        # blocks r = 0 .. C_minus-1 have K_minus bits, the rest K_plus (5.1.2)
        self.groups = ([(s["K_minus"], s["C_minus"])] if s["C_minus"] else []) + [(s["K_plus"], s["C_plus"])]
        kw = dict(n_iter=n_iter, algo=algo, device=device, max_batch=max_batch)
        if algo in ("maxlog_s16", "logmap_s16") and early_term is True:
            # every code block ends in a CRC: its own CRC24B, or the transport block's CRC24A when there is one block
            # (leading filler zeros do not change a CRC): stop each block as soon as its decisions divide
            kw["early_term"] = "crc24b" if s["C"] > 1 else "crc24a"
        elif algo != "logmap_f64":
            kw["early_term"] = True if early_term == "hda" else early_term
        self.dec = {K: TurboDecoder(K, **kw) for K, _ in self.groups}
        self._any = next(iter(self.dec.values()))

    def segment(self, payload):
        """payload uint8 [n_tb, A] on the device -> [(K, blocks [n_tb * C_K, K])], transport-block major."""
        s, n_tb = self.seg, int(payload.shape[0])
        assert payload.is_cuda and tuple(payload.shape) == (n_tb, self.A)
        tb = torch.zeros((n_tb, self.B), dtype=torch.uint8, device=payload.device)
        tb[:, :self.A] = payload
        self._any.crc24_attach(tb, CRC24A)
        if s["C"] == 1:                                         # one block, no CRC24B, fillers in front
            K = s["K_plus"]
            blk = torch.zeros((n_tb, K), dtype=torch.uint8, device=payload.device)
            blk[:, s["F"]:] = tb
            return [(K, blk)]
        # fillers + transport block, cut into the payload parts of the blocks (K - 24 bits each): pure reshapes
        stream = torch.cat([torch.zeros((n_tb, s["F"]), dtype=torch.uint8, device=payload.device), tb], dim=1)
        out, pos = [], 0
        for K, cnt in self.groups:
            blk = torch.zeros((n_tb, cnt, K), dtype=torch.uint8, device=payload.device)
            blk[:, :, :K - 24] = stream[:, pos:pos + cnt * (K - 24)].reshape(n_tb, cnt, K - 24)
            pos += cnt * (K - 24)
            blk = blk.reshape(n_tb * cnt, K)
            self.dec[K].crc24_attach(blk, CRC24B)
            out.append((K, blk))
        return out

    def encode(self, blocks):
        return [(K, self.dec[K].encode(b)) for K, b in blocks]

    def decode(self, llrs):
        """llrs [(K, [n_tb * C_K, 3K+12])] -> (payload [n_tb, A], tb_ok [n_tb], cb_ok [n_tb, C])."""
        s = self.seg
        bits = [(K, self.dec[K].decode(l)["bits"]) for K, l in llrs]
        n_tb = int(bits[-1][1].shape[0]) // self.groups[-1][1]
        dev = bits[0][1].device
        if s["C"] == 1:
            tb = bits[0][1][:, s["F"]:].contiguous()
            cb_ok = torch.ones((n_tb, 1), dtype=torch.uint8, device=dev)
        else:
            oks, parts = [], []
            for (K, cnt), (_, b) in zip(self.groups, bits):
                oks.append(self.dec[K].crc24_check(b, CRC24B).reshape(n_tb, cnt))
                parts.append(b.reshape(n_tb, cnt, K)[:, :, :K - 24].reshape(n_tb, cnt * (K - 24)))
            tb = torch.cat(parts, dim=1)[:, s["F"]:].contiguous()
            cb_ok = torch.cat(oks, dim=1)
        tb_ok = self._any.crc24_check(tb, CRC24A)
        return tb[:, :self.A].contiguous(), tb_ok, cb_ok

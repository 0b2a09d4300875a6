/*
 * turbo_oracle_crc.c -- CPU oracle for the transport-block stage above the decode path
 * (SURVEY.md 8f.3): the LTE 24-bit CRCs and code-block segmentation of TS 36.212 5.1.1 / 5.1.2.
 * TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
 *
 * The reference has only a placeholder for this stage (previous/Decoder.cc:1026 "stoprule ... 1=CRC",
 * :1098-1099 "check stop rule / expected"), so there is nothing to restate from it.
 * Parity status: the CRCs are PINNED to the published check values of the CRC catalogue
 * (CRC-24/LTE-A: poly 0x864CFB, init 0, check("123456789") = 0xCDE703; CRC-24/LTE-B: poly 0x800063,
 * init 0, check = 0x23EF52), tests/test_oracle.py; segmentation is a literal restatement of 5.1.2, UNPINNED.
 */
#include <stdint.h>

#include "turbo_oracle.h"

/* bit-serial CRC, MSB first, zero initial state: the 24 parity bits of bits[0..n) */
unsigned tdo_crc24(const uint8_t *bits, int n, unsigned poly)
{
    unsigned c = 0;
    for (int i = 0; i < n; i++) {
        const unsigned fb = ((c >> 23) & 1u) ^ (bits[i] & 1u);
        c = (c << 1) & 0xffffffu;
        if (fb) c ^= poly & 0xffffffu;
    }
    return c;
}

/* 5.1.2: B = transport block size including its CRC24A.  out = {C, K_plus, K_minus, C_plus, C_minus, F, L} */
int tdo_segmentation(int B, int *out)
{
    const int Z = 6144;
    if (B <= 0) return -1;
    int L = 0, C = 1, Bp = B;
    if (B > Z) { L = 24; C = (B + (Z - L) - 1) / (Z - L); Bp = B + C * L; }
    int Kp = -1, Km = 0;
    for (int i = 0; i < tdo_lte_num_sizes(); i++) {
        const int K = tdo_lte_size_at(i);
        if ((long)C * K >= Bp) { Kp = K; break; }
        Km = K;
    }
    if (Kp < 0) return -1;
    int Cp = C, Cm = 0;
    if (C == 1) Km = 0;
    else {
        const int dK = Kp - Km;
        Cm = (C * Kp - Bp) / dK;
        Cp = C - Cm;
    }
    out[0] = C; out[1] = Kp; out[2] = Km; out[3] = Cp; out[4] = Cm;
    out[5] = Cp * Kp + Cm * Km - Bp; out[6] = L;
    return 0;
}

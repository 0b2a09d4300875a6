"""Extract the reference's published block-error tables into a small fixture.

Run in the dev container only (needs /root/reference):   python tests/golden/make_result_golden.py
Source: ITTC/result.txt of the reference -- run 1 (:2-31, N_ITERATION = 8, at most 10^4 frames per
point) and run 3 (:76-119, N_ITERATION = 15, at most 10^5 frames per point).  Rows = iteration,
columns = Eb/N0 0.0 ... 1.0 dB.  A point stops after 50 block errors at the LAST iteration
(ITTC/main.cpp:239-243), which is how the frame counts are recovered.
"""
import json
import os

SRC = "/root/reference/ITTC/result.txt"
HERE = os.path.dirname(os.path.abspath(__file__))


def table(lines, start):
    rows = []
    for ln in lines[start:]:
        v = ln.split()
        if not v:
            break
        try:
            rows.append([float(x) for x in v])
        except ValueError:
            break
    return rows


def main():
    lines = open(SRC, encoding="latin-1").read().splitlines()
    blers = [i for i, ln in enumerate(lines) if ln.strip() == "Bler:"]
    out = {"source": "ITTC/result.txt", "ebn0_db": [round(0.1 * k, 1) for k in range(11)], "runs": []}
    for idx, max_frames in ((blers[0], 10000), (blers[2], 100000)):
        t = table(lines, idx + 1)
        last = t[-1]
        frames = [max_frames if p * max_frames < 50 - 1e-6 else int(round(50.0 / p)) for p in last]
        out["runs"].append({"lines": "%d-%d" % (idx + 2, idx + 1 + len(t)), "n_iteration": len(t), "max_frames": max_frames,
                            "frames": frames, "bler": t})
    with open(os.path.join(HERE, "ittc_result_bler.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in out["runs"]:
        print(r["lines"], r["n_iteration"], r["frames"])


if __name__ == "__main__":
    main()

"""Regenerates tests/golden/* from the reference's own artefacts.  Run in the build container only
(needs /root/reference); the outputs are committed so tests never read /root/reference.

  faiss_index.bin / .mapping   byte copies of /root/reference/data/faiss_index.bin{,.mapping}
                               (a real FAISS-written IndexFlatL2 file: pins the on-disk layout)
  fixture_answers.json         float64 brute-force answers on that file (numpy only, no oracle code):
                               self-queries + seeded perturbed queries, L2 and IP, k=5 and k=30 (> ntotal)
  synth_known.json             first values of the counter-based synthetic generator (pure-Python ints)
"""
import hashlib
import json
import os
import pickle
import shutil
import struct

import numpy as np

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX = float(np.finfo(np.float32).max)


def brute(x64, q64, k, metric):
    if metric == 1:
        key = ((x64[None, :, :] - q64[:, None, :]) ** 2).sum(2)
    else:
        key = -(q64 @ x64.T)
    out_i, out_d = [], []
    for r in range(q64.shape[0]):
        order = np.lexsort((np.arange(x64.shape[0]), key[r]))[:k]
        ids = order.tolist() + [-1] * (k - len(order))
        ds = [float(key[r, j]) if metric == 1 else float(-key[r, j]) for j in order]
        ds += [FLT_MAX if metric == 1 else -FLT_MAX] * (k - len(order))
        out_i.append(ids)
        out_d.append(ds)
    return out_i, out_d


def mix64(z):
    M = (1 << 64) - 1
    z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & M
    z ^= z >> 27; z = (z * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def synth_int(seed, idx):
    M = (1 << 64) - 1
    a = mix64((seed + 0x9E3779B97F4A7C15 * (2 * idx + 1)) & M)
    b = mix64(a ^ 0xD1B54A32D192ED03)
    c = mix64((b + idx) & M)
    s = 0
    for w in (a, b, c):
        for i in range(4):
            s += (w >> (16 * i)) & 0xFFFF
    return s - 393210


def main():
    for name in ("faiss_index.bin", "faiss_index.bin.mapping"):
        shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, name))
    raw = open(os.path.join(HERE, "faiss_index.bin"), "rb").read()
    assert raw[:4] == b"IxF2"
    d, n = struct.unpack("<iq", raw[4:16])
    x = np.frombuffer(raw, "<f4", n * d, 45).reshape(n, d)
    mapping = pickle.load(open(os.path.join(HERE, "faiss_index.bin.mapping"), "rb"))
    x64 = x.astype(np.float64)
    rng = np.random.default_rng(20261018)
    pert = (x[rng.integers(0, n, 16)] + rng.standard_normal((16, d)).astype(np.float32) * 0.5).astype(np.float32)
    cases = []
    for metric in (1, 0):
        for k in (5, 30):
            for qname, q in (("self", x), ("perturbed", pert)):
                ids, ds = brute(x64, q.astype(np.float64), k, metric)
                cases.append({"metric": metric, "k": k, "queries": qname, "ids": ids, "dists": ds})
    ans = {
        "sha256_index": hashlib.sha256(raw).hexdigest(),
        "sha256_mapping": hashlib.sha256(open(os.path.join(HERE, "faiss_index.bin.mapping"), "rb").read()).hexdigest(),
        "d": d, "ntotal": n, "mapping": mapping,
        "perturbed_queries": pert.astype(np.float64).tolist(),
        "cases": cases,
    }
    json.dump(ans, open(os.path.join(HERE, "fixture_answers.json"), "w"))
    known = {"seed": 1234, "d": 384,
             "ints_row0": [synth_int(1234, i) for i in range(16)],
             "ints_row1000003": [synth_int(1234, 1000003 * 384 + i) for i in range(16)],
             "seed_q": 5678, "ints_q_row7": [synth_int(5678, 7 * 384 + i) for i in range(16)]}
    json.dump(known, open(os.path.join(HERE, "synth_known.json"), "w"))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()

"""Small run of the general (W, O) kernel for compute-sanitizer: pairs with exhausted texts and windows that end exactly at
the end of the packed blobs, and mapping candidates at the very end of the genome, at four window configurations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scrooge_b200  # noqa: E402
from conftest import random_pairs  # noqa: E402

for W, O in ((128, 65), (96, 49), (48, 25), (64, 20), (16, 9)):
    al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", f"golden_w{W}_o{O}.json")))
    T = [x["text"] for v in g["groups"].values() for x in v]
    Q = [x["query"] for v in g["groups"].values() for x in v]
    T2, Q2 = random_pairs(5 + W, 300, [0, 1, W - 1, W, W + 1, 150, 700], [0, 0.1, 0.4])
    res = al.align_pairs(T + T2, Q + Q2)
    want = [x["cigar"] for v in g["groups"].values() for x in v]
    assert res.cigars()[: len(want)] == want, (W, O)
    m = g["mapping"]
    cs = [s for l in m["locations"] for s in l]
    cr = [r for r, l in enumerate(m["locations"]) for _ in l]
    al.set_reference(m["genome"])
    assert al.align_candidates(m["reads"], cs, cr).cigars() == m["cigar"], (W, O)
    al.close()
    print("ok", W, O, flush=True)

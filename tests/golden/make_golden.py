"""Generates the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src/genasm_cpu.cpp).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs: the reference's own known-answer reads (src/tests.cu:236-244), its 8 differential pairs
(src/tests.cu:275-284), the library example (src/library_example.cu:12-13), hand-written edge cases
 and seeded random pairs.  Outputs per (W, O) build: edit distance + CIGAR string.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.binding import CONFIGS, EXTRA_CONFIGS, RefCpu  # noqa: E402

KAT_REF = "AAAACCCCGGGGTTTT"
KAT_READS = [
    "CCCCGGGGTTTTAAAA", "AAAACCCCGGGGTTTT", "ACCCCGG", "AAAAGGGGAAAATTTT", "AAAAAAAAAAAAAAAA", "ATTAACGCCTTT",
    "TTTTAAAACCCCGGGGTTTTAAAA", "", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTAAAACCCCGGGGTTTTAAAA",
]
KAT_DISTANCES = [8, 0, 3, 8, 12, 6, 8, 0, 48]  # src/tests.cu:246

# (query, text) as in src/tests.cu:275-284
DIFF_PAIRS = [
    ("ACGT", "ACGT"),
    ("CAAATCTATTAAGTCAAACGGTCCGTAAGCTAGAACCTCCTGCCGTGTAAGTTACGACGTGGTCGAGTTACTTTCGTTCTTATTAACACAATGTCCATCA", "CAAACCTATCAAGTCAAACGGTCCGTAGCTACACCTCCTGCCGTGTAAAGTTACGACGTGGTTGAGTTACTTTCGTTCTTATTAACAACAATGTTCCATCA"),
    ("CGGCGAAGGAATTAATTACAAGCCTTGTACACTTGCATATTCTTCTGCAACAGGGCCCCGGCTCCGTCCTACCTCGGTTTACTGTGACTCACTTGAGCGA", "CGGCGAAGGAATAATTACAAGCCTGTATCACTTGCATATTCGTTCTGCAACAGGCCCGGCTCCGTCTACGCTGGTTTACTGTGACTCACTTGAGCGA"),
    ("ACAGTGGAAATGTCGCGGAAGGGTAGCAGTAGAACTTAATCAGAGAGATTACCTCGCGTAGTTGAAGTCTTGACGGGCGCATTGGACATAACAAACATAC", "ACGTGGACATGTCGCGGAAGGATAGCAGTAGAACTTAATCAGAGAATTACCTCGCGTAGTTGAACTCTTGACGGCGCGATGTGGACCTAACAAACATAC"),
    ("AACCCACGGTCTTCTCTGGTTTCGAACTTACAATCGTGAGCCCATCCGTACTTTCATGTTTCTTAAGATGGCAAGACAGAAATATAATTAGGCCGGGAGC", "AACCCACGGTCTTCTCTGGTTTCGAATTAGCAATCGTCGAGCCGCATCCGTACTTTCATGTTTCCTTAAGATGGCCAGAACAGAAATAATTAGGCCGGGAGC"),
    ("TTTGCTTAGCCGAGCTATGCGGAACTAGAGCACCGGAGGTTTGTGTGGTCACTAGAATGACAAGGTCTCTGATCAGATATAACTCTTCGGGTTTGCGTAA", "TTTGCTTAGCCGAGCTATCCCGGAACAGACACCGGAGGTTTGAGTGGTCACTAGAATGACAAGGTATCTGATCAGATACAACTTCTTCGGGCTTTGCGTAA"),
    ("GATGTACAGTCTCGAAAACCAAGTCTAGGACCAATTCCAACCTTATAATCCAGATTTACCATTATGACAACCGCAGAAGAGAAACTAATCGTCCAAAAGA", "GATGTGCAGTCTCGAAAACCAAGTCTAGGACCAGATTCCAACCTTTTAACCCAGAGTTACCAGAGACAACCGCAGAAGAGAAACTAATCGTCCAAAAGA"),
    ("TCCTGCGCGCGAAGGGGACATTGCAGGGCAAAGCAATGGCTAGATAGCCTCATACTGAGACGATAAATGGCGTTGGACACCGGAGAAAAGACCCCGCCGA", "TCTGCGCGCGAAGGGGACATAGCAGGCAAAGCAATGGCTAGATAGCCTCATACTGAGAGATAAATGGCGTTGGCCACCGGAGCAAAAGACCCCGCCG"),
]


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def mutate(rng, t, L, e):
    out, i = [], 0
    while len(out) < L and i < len(t):
        if rng.random() < e:
            k = rng.randrange(3)
            if k == 0:
                out.append(rng.choice([c for c in "ACGT" if c != t[i]])); i += 1
            elif k == 1:
                out.append(rng.choice("ACGT"))
            else:
                i += 1
        else:
            out.append(t[i]); i += 1
    while len(out) < L:
        out.append(rng.choice("ACGT"))
    return "".join(out)


def edge_cases(rng):
    """(text, query) pairs of SURVEY.md section 8c's edge set."""
    cases = [("", ""), ("ACGT", ""), ("", "ACGTACGTAC"), ("A", "A"), ("A", "C"), ("ACGTACGT", "ACGTACG")]
    for L in (1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 66, 67, 127, 128, 129):
        t = rand_seq(rng, L + 70)
        cases.append((t, t[:L]))                       # exact prefix
        cases.append((t, mutate(rng, t, L, 0.1)))     # typical
        cases.append((t[: L // 2], mutate(rng, t, L, 0.1)))  # read >> text: insertion tail (Q5)
    cases.append((rand_seq(rng, 3000), rand_seq(rng, 40)))       # text >> read
    cases.append(("A" * 300, "A" * 200))                          # all the same base
    cases.append(("A" * 300, "C" * 200))                          # nothing matches
    cases.append((rand_seq(rng, 400), rand_seq(rng, 300)))        # unrelated pair, d_w ~ 32
    t = rand_seq(rng, 400)
    cases.append((t.lower(), mutate(rng, t, 300, 0.05)))          # lowercase text
    cases.append((t, mutate(rng, t, 300, 0.05).lower()))          # lowercase query
    return cases


def main():
    # the two configurations the reference ships, then the extra window configurations (smaller sets)
    for W, O in list(CONFIGS.items()) + list(EXTRA_CONFIGS):
        extra = CONFIGS.get(W) != O
        ref = RefCpu(W, O)
        rng = random.Random(1000 + W if not extra else 5000 + 131 * W + O)
        out = {"W": W, "O": O, "generator": "tests/golden/make_golden.py", "groups": {}}

        def run(name, pairs):
            res = ref.align_pairs([p[0] for p in pairs], [p[1] for p in pairs])
            out["groups"][name] = [
                {"text": t, "query": q, "edit": int(res.edit[k]), "cigar": res.cigars[k]} for k, (t, q) in enumerate(pairs)
            ]

        run("kat_tests_cu", [(KAT_REF, r) for r in KAT_READS])
        run("differential_tests_cu", [(t, q) for (q, t) in DIFF_PAIRS])
        run("library_example", [("ACGTACGT", "ACGTACG")])
        run("edge", edge_cases(rng))
        rnd = []
        for _ in range(30 if extra else 120):
            L = rng.choice([20, 50, 100, 150, 250, 400, 1000])
            e = rng.choice([0.0, 0.02, 0.05, 0.1, 0.15, 0.3])
            t = rand_seq(rng, L + L // 3 + 64)
            rnd.append((t, mutate(rng, t, L, e)))
        run("random", rnd)
        if W == 64 and not extra:
            assert [g["edit"] for g in out["groups"]["kat_tests_cu"]] == KAT_DISTANCES
        # mapping interface: one genome, reads with several candidate starts (unaligned mod 4, genome end)
        genome = rand_seq(rng, 5000)
        reads, locs = [], []
        for r in range(12):
            L = rng.choice([30, 100, 150, 600])
            s = rng.randrange(0, len(genome) - L)
            reads.append(mutate(rng, genome[s:], L, 0.08))
            locs.append([s, max(0, s - rng.randrange(1, 9)), min(len(genome), s + rng.randrange(1, 9))])
        reads.append(mutate(rng, genome[4990:], 40, 0.0))
        locs.append([4990, 4999, 5000])  # candidates at the very end of the genome
        res = ref.align_mapping(genome, reads, locs)
        out["mapping"] = {"genome": genome, "reads": reads, "locations": locs,
                          "edit": [int(x) for x in res.edit], "cigar": res.cigars}
        with open(os.path.join(HERE, f"golden_w{W}_o{O}.json" if extra else f"golden_w{W}.json"), "w") as f:
            json.dump(out, f, indent=0)
        print("wrote", f.name, sum(len(v) for v in out["groups"].values()), "pairs +", len(res.cigars), "candidates")


if __name__ == "__main__":
    main()

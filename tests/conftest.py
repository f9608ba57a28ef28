import json
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    out = {}
    for W in (64, 32):
        with open(os.path.join(ROOT, "tests", "golden", f"golden_w{W}.json")) as f:
            out[W] = json.load(f)
    return out


@pytest.fixture(scope="session")
def golden_wo():
    """Golden vectors of the extra window configurations, keyed by (W, O) (oracle/binding.py:EXTRA_CONFIGS)."""
    from oracle.binding import EXTRA_CONFIGS
    out = {}
    for W, O in EXTRA_CONFIGS:
        with open(os.path.join(ROOT, "tests", "golden", f"golden_w{W}_o{O}.json")) as f:
            out[(W, O)] = json.load(f)
        assert out[(W, O)]["W"] == W and out[(W, O)]["O"] == O
    return out


@pytest.fixture(scope="session")
def sglib():
    import scrooge_b200
    if not os.path.exists(scrooge_b200._lib.LIB_PATH):
        scrooge_b200.build()
    return scrooge_b200.lib()


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


def random_pairs(seed, count, lengths, errors, unrelated=0.1, short_text=0.2):
    """Mixed bag: mutated reads, unrelated reads, texts cut short (insertion tails), empty strings."""
    rng = random.Random(seed)
    T, Q = [], []
    for _ in range(count):
        L = rng.choice(lengths)
        t = rand_seq(rng, L + L // 3 + 70)
        q = rand_seq(rng, L) if rng.random() < unrelated else mutate(rng, t, L, rng.choice(errors))
        if rng.random() < short_text:
            t = t[: rng.choice([0, 1, L // 2, L])]
        T.append(t)
        Q.append(q)
    return T, Q


def mapping_case(seed, n_reads=300, genome_len=60000):
    """A small read-mapping case: per read the true locus, two jittered ones and an unrelated one (candidates may run into
    the end of the genome).  Returns (genome, reads, cand_start, cand_read)."""
    rng = random.Random(seed)
    genome = rand_seq(rng, genome_len)
    reads, cs, cr = [], [], []
    for r in range(n_reads):
        L = rng.choice([1, 40, 150, 900])
        p0 = rng.randrange(0, len(genome) - L - 10)
        rd = list(genome[p0:p0 + L])
        for k in range(len(rd)):
            if rng.random() < 0.08:
                rd[k] = rng.choice("ACGT")
        reads.append("".join(rd))
        for c in (p0, max(0, p0 - 3), min(len(genome) - 1, p0 + 5), rng.randrange(0, len(genome))):
            cs.append(c)
            cr.append(r)
    return genome, reads, cs, cr

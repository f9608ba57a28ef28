"""The algorithmic identities the CUDA kernel relies on (left-aligned vectors, G-row chunks, V/H/E edge words
replacing stored R rows -- see scrooge_b200/csrc/sg_align.cuh) checked on the CPU against the oracle."""
import pytest

from conftest import random_pairs
from kernel_model import align


@pytest.mark.parametrize("W,O", [(64, 33), (32, 17)])
def test_model_matches_oracle(oracle, W, O):
    T, Q = random_pairs(11 + W, 150, [0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 200], [0, 0.05, 0.15, 0.4, 0.8])
    res = oracle.align_pairs(T, Q, W=W)
    for k in range(len(T)):
        ed, cg, rc = align(T[k], Q[k], W, O, G=8 if W == 64 else 4)  # the kernel's chunk sizes
        assert (ed, cg, rc) == (int(res.edit[k]), res.cigars[k], int(res.ref_consumed[k])), (T[k], Q[k])


def test_model_golden_kats(golden):
    for x in golden[64]["groups"]["kat_tests_cu"] + golden[64]["groups"]["differential_tests_cu"]:
        ed, cg, _ = align(x["text"], x["query"])
        assert ed == x["edit"] and cg == x["cigar"]


@pytest.mark.parametrize("W,O", [(64, 33), (32, 17)])
def test_delta_model_matches_oracle(oracle, W, O):
    """The delta-encoded (column-wise +-1) DC yields the same distances, CIGARs, consumed prefixes and the same
    early-termination-minimal entry count as the reference's row-wise threshold vectors."""
    from kernel_model import align_delta
    T, Q = random_pairs(23 + W, 300, [0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 200, 500], [0, 0.05, 0.15, 0.4, 0.8])
    res = oracle.align_pairs(T, Q, W=W)
    total = 0
    for k in range(len(T)):
        ed, cg, rc, ent = align_delta(T[k], Q[k], W, O)
        total += ent
        assert (ed, cg, rc) == (int(res.edit[k]), res.cigars[k], int(res.ref_consumed[k])), (T[k], Q[k])
    assert total == res.stats["dc_entries"]


def test_delta_model_golden_kats(golden):
    from kernel_model import align_delta
    for W in (64, 32):
        for x in golden[W]["groups"]["kat_tests_cu"] + golden[W]["groups"]["differential_tests_cu"]:
            ed, cg, _, _ = align_delta(x["text"], x["query"], W, 33 if W == 64 else 17)
            assert ed == x["edit"] and cg == x["cigar"]


def _generic_configs():
    from oracle.binding import CONFIGS, EXTRA_CONFIGS
    return list(CONFIGS.items()) + list(EXTRA_CONFIGS)


@pytest.mark.parametrize("W,O", _generic_configs())
def test_generic_delta_model_matches_oracle(oracle, W, O):
    """The run-time (W, O) kernel's formulation -- vectors padded to a multiple of 32 bits, W columns per window whatever
    n is, full-width op planes, any traceback limit -- against the oracle on every window configuration."""
    from kernel_model import align_delta_generic
    T, Q = random_pairs(41 + 7 * W + O, 120, [0, 1, 2, W // 2, W - 1, W, W + 1, 2 * W + 1, 100, 300], [0, 0.05, 0.15, 0.4, 0.8])
    res = oracle.align_pairs(T, Q, W=W, O=O)
    total = 0
    for k in range(len(T)):
        ed, cg, rc, ent = align_delta_generic(T[k], Q[k], W, O)
        total += ent
        assert (ed, cg, rc) == (int(res.edit[k]), res.cigars[k], int(res.ref_consumed[k])), (W, O, T[k], Q[k])
    assert total == res.stats["dc_entries"]


@pytest.mark.parametrize("W,O", [(2, 1), (3, 1), (5, 0), (33, 2), (63, 0), (65, 3), (100, 40), (127, 64), (96, 16), (127, 3),
                                 (100, 0), (128, 64), (128, 0), (192, 97), (255, 127), (200, 72), (256, 250)])
def test_odd_windows_model(oracle, W, O):
    """Tiny windows and window sizes that are not a multiple of 32: the general kernel's formulation against the oracle."""
    from kernel_model import align_delta_generic
    T, Q = random_pairs(51 + 7 * W + O, 60, [0, 1, 2, W - 1, W, W + 1, 2 * W + 1, 150], [0, 0.1, 0.4])
    res = oracle.align_pairs(T, Q, W=W, O=O)
    for k in range(len(T)):
        ed, cg, rc, _ = align_delta_generic(T[k], Q[k], W, O)
        assert (ed, cg, rc) == (int(res.edit[k]), res.cigars[k], int(res.ref_consumed[k])), (W, O, T[k], Q[k])


def test_random_windows_model(oracle):
    """Thirty random (W, O) of the supported range: the general kernel's formulation (padded vectors, full-width planes, op
    streams of 4 or 8 words, stream RLE) against the oracle."""
    import random
    from kernel_model import align_delta_generic
    rng = random.Random(4711)
    for _ in range(30):
        W = rng.choice([rng.randint(2, 256), rng.choice([31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256])])
        O = rng.randint(max(0, W - 128), W - 1)
        T, Q = random_pairs(rng.randrange(1 << 30), 16, [0, 1, 2, W - 1, W, W + 1, 2 * W + 1, 300], [0, 0.05, 0.3, 0.6])
        res = oracle.align_pairs(T, Q, W=W, O=O)
        for k in range(len(T)):
            ed, cg, rc, _ = align_delta_generic(T[k], Q[k], W, O)
            assert (ed, cg, rc) == (int(res.edit[k]), res.cigars[k], int(res.ref_consumed[k])), (W, O, T[k], Q[k])


def test_unchecked_walk_with_cut_equals_checked_walk():
    """The short-window walk of genasm_delta_kernel: TB_LIMIT unchecked steps + cutting the op streams where the pattern
    ran out gives the checked walk's ops, i and j -- for ANY plane contents (random planes: whatever lies below the
    pattern cannot matter), every m, both tuned window configurations."""
    import random
    from kernel_model import walk_checked, walk_unchecked_then_cut
    rng = random.Random(2024)
    for TBL in (31, 15):
        for trial in range(4000):
            m = rng.randrange(1, TBL + 8)
            bias = rng.choice([0.02, 0.2, 0.5, 0.9])     # from almost all '=' to mostly edits (long walks, many 'D')
            A = [sum((rng.random() < bias) << b for b in range(32)) for _ in range(TBL + 1)]
            B = [rng.getrandbits(32) for _ in range(TBL + 1)]
            assert walk_unchecked_then_cut(A, B, m, TBL) == walk_checked(A, B, m, TBL), (TBL, m, trial)

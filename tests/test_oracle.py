"""The oracle (oracle/genasm_oracle.c) against the reference's own known answers, the committed golden
vectors made from the unmodified reference, and -- when oracle/_ref is present -- the reference itself."""
import pytest

from conftest import random_pairs
from oracle.binding import EXTRA_CONFIGS, RefCpu, cigar_ref_consumed

KAT_REF = "AAAACCCCGGGGTTTT"
KAT_DIST = [8, 0, 3, 8, 12, 6, 8, 0, 48]  # reference src/tests.cu:246


def test_kat_distances_and_validator(oracle, golden):
    g = golden[64]["groups"]["kat_tests_cu"]
    res = oracle.align_pairs([x["text"] for x in g], [x["query"] for x in g])
    assert list(res.edit) == KAT_DIST
    assert res.cigars[0] == "4D12=4I" and res.cigars[8] == "31I13I16=2I2I" and res.cigars[7] == ""
    for x, cg, ed in zip(g, res.cigars, res.edit):
        assert oracle.validate_cigar(cg, x["text"], x["query"], int(ed)) == 0


def test_library_example(oracle):
    res = oracle.align_pairs(["ACGTACGT"], ["ACGTACG"])  # reference src/library_example.cu:12-13
    assert int(res.edit[0]) == 0 and res.cigars[0] == "7="


@pytest.mark.parametrize("W", [64, 32])
def test_golden_groups(oracle, golden, W):
    for name, g in golden[W]["groups"].items():
        res = oracle.align_pairs([x["text"] for x in g], [x["query"] for x in g], W=W)
        for k, x in enumerate(g):
            assert int(res.edit[k]) == x["edit"], (name, k)
            assert res.cigars[k] == x["cigar"], (name, k)
            assert int(res.ref_consumed[k]) == cigar_ref_consumed(x["cigar"])
            assert oracle.validate_cigar(res.cigars[k], x["text"], x["query"], x["edit"]) == 0


@pytest.mark.parametrize("W", [64, 32])
def test_golden_mapping(oracle, golden, W):
    m = golden[W]["mapping"]
    cs = [s for l in m["locations"] for s in l]
    cr = [r for r, l in enumerate(m["locations"]) for _ in l]
    res = oracle.align_candidates(m["genome"], m["reads"], cs, cr, W=W)
    assert [int(x) for x in res.edit] == m["edit"]
    assert res.cigars == m["cigar"]
    for c in range(len(cs)):
        assert oracle.validate_cigar(res.cigars[c], m["genome"], m["reads"][cr[c]], m["edit"][c], ref_start=cs[c]) == 0


@pytest.mark.parametrize("W", [64, 32])
def test_live_reference(oracle, W):
    if not RefCpu.available(W):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    ref = RefCpu(W)
    T, Q = random_pairs(77 + W, 1500, [0, 1, 3, 33, 64, 65, 100, 150, 700], [0, 0.03, 0.1, 0.2, 0.5])
    a = oracle.align_pairs(T, Q, W=W, threads=2)
    b = ref.align_pairs(T, Q, threads=2)
    assert list(a.edit) == list(b.edit)
    assert a.cigars == b.cigars


@pytest.mark.parametrize("W,O", EXTRA_CONFIGS)
def test_golden_window_configurations(oracle, golden_wo, W, O):
    """The window sweep's configurations (scripts/profile.py:66-100,595-640): goldens from the unmodified reference built
    with -DCLI_W=<W> -DCLI_K=<W> -DCLI_O=<O>, W up to 128 (multi-word vectors, src/bitvector.hpp:45-47)."""
    gw = golden_wo[(W, O)]
    for name, g in gw["groups"].items():
        res = oracle.align_pairs([x["text"] for x in g], [x["query"] for x in g], W=W, O=O)
        for k, x in enumerate(g):
            assert int(res.edit[k]) == x["edit"], (name, k)
            assert res.cigars[k] == x["cigar"], (name, k)
            assert int(res.ref_consumed[k]) == cigar_ref_consumed(x["cigar"])
            assert oracle.validate_cigar(res.cigars[k], x["text"], x["query"], x["edit"]) == 0
    m = gw["mapping"]
    cs = [s for l in m["locations"] for s in l]
    cr = [r for r, l in enumerate(m["locations"]) for _ in l]
    res = oracle.align_candidates(m["genome"], m["reads"], cs, cr, W=W, O=O)
    assert [int(x) for x in res.edit] == m["edit"] and res.cigars == m["cigar"]


@pytest.mark.parametrize("W,O", EXTRA_CONFIGS)
def test_live_reference_window_configurations(oracle, W, O):
    if not RefCpu.available(W, O):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    ref = RefCpu(W, O)
    T, Q = random_pairs(177 + 131 * W + O, 500, [0, 1, 3, W - 1, W, W + 1, 2 * W + 3, 150, 700], [0, 0.03, 0.1, 0.2, 0.5])
    a = oracle.align_pairs(T, Q, W=W, O=O, threads=2)
    b = ref.align_pairs(T, Q, threads=2)
    assert list(a.edit) == list(b.edit)
    assert a.cigars == b.cigars


def test_window_limits(oracle):
    with pytest.raises(ValueError):
        oracle.align_pairs(["ACGT"], ["ACG"], W=257, O=129)
    with pytest.raises(ValueError):
        oracle.align_pairs(["ACGT"], ["ACG"], W=64, O=64)


def test_threads_do_not_change_results(oracle):
    T, Q = random_pairs(5, 300, [100, 150, 1000], [0.05, 0.1])
    a = oracle.align_pairs(T, Q, threads=1)
    b = oracle.align_pairs(T, Q, threads=4)
    assert list(a.edit) == list(b.edit) and a.cigars == b.cigars and a.stats == b.stats


def test_dc_entry_counter(oracle):
    # one exact 16-base window: d_w = 0, n = 16 -> (0+1)*(16+1) entries
    res = oracle.align_pairs(["ACGTACGTACGTACGT"], ["ACGTACGTACGTACGT"])
    assert res.stats == dict(windows=1, dc_entries=17, tb_steps=16)


def test_twobit_reference_layout(oracle):
    # reference src/tests.cu:583-606: first base of a byte in bits 7:6, tail zero padded
    assert oracle.twobit_ref_layout("") == b""
    assert oracle.twobit_ref_layout("A") == bytes([0x00])
    assert oracle.twobit_ref_layout("T") == bytes([0xC0])
    assert oracle.twobit_ref_layout("ACGT") == bytes([0b00011011])
    assert oracle.twobit_ref_layout("ACGTA") == bytes([0b00011011, 0x00])
    assert oracle.twobit_ref_layout("ACGTC") == bytes([0b00011011, 0x40])
    assert oracle.twobit_ref_layout("acgt" * 8 + "g") == bytes([0b00011011] * 8 + [0x80])


def test_bad_base_is_reported(oracle):
    with pytest.raises(ValueError):
        oracle.align_pairs(["ACGN"], ["ACG"])

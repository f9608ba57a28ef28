"""Dataset readers (include/scrooge_io.hpp) and the sg_tests driver: on the CPU the parsed inputs are checked against
expectations and, when oracle/_ref is built, against the reference's own readers (src/util.cpp); on the GPU the driver
aligns the files and validates every CIGAR like the reference's `tests` binary does."""
import os
import random
import subprocess

import pytest

from conftest import mutate, rand_seq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SG_TESTS = os.path.join(ROOT, "build", "sg_tests")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_io_dump")


def wrap(seq, width):
    return "\n".join(seq[i:i + width] for i in range(0, len(seq), width))


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    """Two chromosomes, 30 reads sampled from them with a clipped prefix (so left extension matters), both strands in
    the seed files, CRLF line ends in the FASTQ."""
    d = tmp_path_factory.mktemp("dataset")
    rng = random.Random(3)
    chroms = {"chrA": rand_seq(rng, 6000), "chrB test": rand_seq(rng, 4000)}
    with open(d / "ref.fasta", "w") as f:
        for name, seq in chroms.items():
            f.write(f">{name}\n{wrap(seq, 70)}\n")
    starts, off = {}, 0
    for name, seq in chroms.items():
        starts[name] = off
        off += len(seq)
    reads, seeds = [], []
    for r in range(30):
        name = rng.choice(list(chroms))
        L = rng.choice([60, 150, 400, 1500])
        s = rng.randrange(0, len(chroms[name]) - L - 200)
        content = mutate(rng, chroms[name][s:], L, 0.08)
        clip = rng.randrange(0, 20)
        strand = "+" if rng.random() < 0.8 else "-"
        reads.append((f"read_{r}", content))
        seeds.append(dict(read=f"read_{r}", qlen=L, qstart=clip, qend=L - 3, strand=strand, chrom=name, tlen=len(chroms[name]),
                          tstart=s + clip, tend=s + L))
    with open(d / "reads.fastq", "w", newline="") as f:
        for name, content in reads:
            f.write(f"@{name}\r\n{content}\r\n+\r\n{'I' * len(content)}\r\n")
    with open(d / "seeds.paf", "w") as f:
        for x in seeds:
            f.write("\t".join(map(str, [x["read"], x["qlen"], x["qstart"], x["qend"], x["strand"], x["chrom"], x["tlen"], x["tstart"],
                                        x["tend"], x["qend"] - x["qstart"], x["qend"] - x["qstart"], 60])) + "\n")
    # MAF is single-chromosome in the reference's reader (the ref line's source must be "ref")
    with open(d / "ref1.fasta", "w") as f:
        f.write(f">only\n{wrap(chroms['chrA'], 60)}\n")
    with open(d / "seeds.maf", "w") as f:
        f.write("##maf version=1\n\n")
        for x in seeds:
            if x["chrom"] != "chrA":
                continue
            f.write("a\n")
            f.write(f"s ref {x['tstart']} {x['tend'] - x['tstart']} + {x['tlen']} ACGT\n")
            f.write(f"s {x['read']} {x['qstart']} {x['qend'] - x['qstart']} {x['strand']} {x['qlen']} ACGT\n\n")
    return dict(dir=d, chroms=chroms, starts=starts, reads=dict(reads), seeds=seeds)


def dump(binary, *files):
    r = subprocess.run([binary, *files] if binary == REF_DUMP else
                       [binary, "--dump_inputs", f"--reference={files[0]}", f"--reads={files[1]}", f"--seeds={files[2]}"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_paf_dataset_is_parsed(sglib, dataset):
    d = dataset["dir"]
    out = dump(SG_TESTS, d / "ref.fasta", d / "reads.fastq", d / "seeds.paf")
    lines = out.strip().splitlines()
    assert lines[0] == "genome 10000 bases, 2 chromosome(s)"
    assert 'chromosome "chrA" starts at 0' in lines and 'chromosome "chrB test" starts at 6000' in lines
    parsed = {}
    for l in lines:
        if l.startswith("read "):
            name = l.split('"')[1]
            rest = l.split('"')[2].split()
            parsed[name] = (int(rest[0]), " ".join(rest[2:]))
    assert len(parsed) == 30
    lens = [int(l.split('"')[2].split()[0]) for l in lines if l.startswith("read ")]
    assert lens == sorted(lens, reverse=True)  # longest first, as the reference's driver sorts them
    for x in dataset["seeds"]:
        n, locs = parsed[x["read"]]
        assert n == len(dataset["reads"][x["read"]])
        if x["strand"] == "-":
            assert locs == ""  # reverse-strand candidates are dropped
            continue
        start = x["tstart"] - x["qstart"]  # left extension to the read's first base
        assert locs == f"{x['chrom']}@{start}->{dataset['starts'][x['chrom']] + start}"


@pytest.mark.parametrize("seeds,ref", [("seeds.paf", "ref.fasta"), ("seeds.maf", "ref1.fasta")])
def test_readers_match_the_reference(sglib, dataset, seeds, ref):
    if not os.path.exists(REF_DUMP):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    d = dataset["dir"]
    files = (d / ref, d / "reads.fastq", d / seeds)
    assert dump(SG_TESTS, *files) == dump(REF_DUMP, *files)


def test_option_parsing(sglib):
    r = subprocess.run([SG_TESTS, "--nonsense"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("sg_tests [options]")
    r = subprocess.run([SG_TESTS, "--reference"], capture_output=True, text=True)  # a valued option without a value
    assert r.stdout.startswith("sg_tests [options]")
    r = subprocess.run([SG_TESTS, "--gpu_info_only"], capture_output=True, text=True)
    assert "visible GPU(s)" in r.stdout


@pytest.mark.gpu
def test_unit_tests_binary(sglib):
    r = subprocess.run([SG_TESTS, "--unit_tests"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for t in ("ascii_to_two_bit_correctness_test", "gpu_algorithm_correctness_test", "library_interface_correctness_test"):
        assert f"PASSED {t}" in r.stdout, r.stdout
    assert "FAILED" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("seeds,ref", [("seeds.paf", "ref.fasta"), ("seeds.maf", "ref1.fasta")])
def test_performance_test_on_files(sglib, dataset, seeds, ref):
    d = dataset["dir"]
    r = subprocess.run([SG_TESTS, f"--reference={d / ref}", f"--reads={d / 'reads.fastq'}", f"--seeds={d / seeds}"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed the sanity check" in r.stdout and "GPU kernel ran at" in r.stdout and "FAILED" not in r.stdout
    n = sum(1 for x in dataset["seeds"] if x["strand"] == "+" and (ref == "ref.fasta" or x["chrom"] == "chrA"))
    assert f"{n} alignments," in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("window,overlap", [(128, 65), (48, 25), (64, 0), (256, 129)])
def test_performance_test_on_files_other_windows(sglib, dataset, window, overlap):
    """The reference's file-based workflow at other window configurations (a rebuild there, --window/--overlap here): every
    CIGAR the C++ drop-in returns passes the validateCigarString port."""
    d = dataset["dir"]
    r = subprocess.run([SG_TESTS, f"--reference={d / 'ref.fasta'}", f"--reads={d / 'reads.fastq'}", f"--seeds={d / 'seeds.paf'}",
                        f"--window={window}", f"--overlap={overlap}"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed the sanity check" in r.stdout and "FAILED" not in r.stdout

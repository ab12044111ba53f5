"""specimine: the product's specimine must write the same `.mined` file as the UNMODIFIED
reference's (goldens: oracle/make_specimine_goldens.py -> tests/golden_specimine/cases.json.gz).  The CPU tier feeds
the module distances from the oracle's aligner (test-only injection) so that file discovery, the identity rule and the
output naming are checked without a GPU; the GPU tier runs the real thing through smx_hw_distances."""
import gzip
import json
import os

import numpy as np
import pytest

import helpers as H
from specimux_b200 import specimine


def _cases():
    with gzip.open(os.path.join(H.HERE, "golden_specimine", "cases.json.gz"), "rt") as fh:
        return json.load(fh)


def _materialise(case, root):
    for rel, text in case["files"].items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as fh:
            fh.write(text)
    return ["--index", os.path.join(root, "index.txt"), "--fastq", os.path.join(root, case["target"])] + case["flags"]


def _oracle_distances(patterns, max_dist, texts, device=0):
    from oracle import aligner
    out = np.empty((len(patterns), len(texts)), dtype=np.int32)
    for i, (p, k) in enumerate(zip(patterns, max_dist)):
        for j, t in enumerate(texts):
            out[i, j] = aligner.align(p, t, "HW", "distance", k, None)["editDistance"]
    return out


@pytest.mark.parametrize("idx", range(5))
def test_mined_file_equals_reference_with_oracle_distances(idx, tmp_path, monkeypatch):
    case = _cases()[idx]
    argv = _materialise(case, str(tmp_path))
    monkeypatch.setattr(specimine, "hw_distances", _oracle_distances)
    specimine.main(argv)
    with open(os.path.join(str(tmp_path), case["target"] + ".mined")) as fh:
        assert fh.read() == case["mined"]


def test_helpers_follow_the_reference_rules(tmp_path):
    assert specimine.extract_specimen_id("/x/full/ITS/sample_AB-1.fastq") == "AB-1"
    assert specimine.extract_specimen_id("SP.2.fastq") == "SP.2"
    with pytest.raises(ValueError):
        specimine.extract_specimen_id("reads.fq")
    assert specimine.detect_input_level("/o/full/ITS/SP1.fastq") == ("/o", "ITS", None)
    assert specimine.detect_input_level("/o/full/ITS/A-B/SP1.fastq") == ("/o", "ITS", "A-B")
    with pytest.raises(ValueError):
        specimine.detect_input_level("/o/partial/ITS/SP1.fastq")

    class R:
        pass
    r = R()
    r.id, r.description = "a_mined_reverse_0.93", "a\tdx:i:0 mined_reverse identity=0.93"
    assert specimine.fastq_title(r) == "a_mined_reverse_0.93 a\tdx:i:0 mined_reverse identity=0.93"
    r.id, r.description = "a", "a extra"
    assert specimine.fastq_title(r) == "a extra"


def test_no_gpu_means_no_distances():
    from specimux_b200 import _lib
    if _lib.load().smx_device_count() > 0:
        pytest.skip("box has a GPU")
    with pytest.raises(_lib.SmxError) as ei:
        specimine.hw_distances(["ACGT"], [1], ["ACGTT"])
    assert ei.value.code == _lib.SMX_ERR_NO_DEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(5))
def test_mined_file_equals_reference_on_gpu(idx, tmp_path):
    case = _cases()[idx]
    argv = _materialise(case, str(tmp_path))
    specimine.main(argv)
    with open(os.path.join(str(tmp_path), case["target"] + ".mined")) as fh:
        assert fh.read() == case["mined"]


@pytest.mark.gpu
def test_hw_distances_match_oracle_on_random_pairs():
    """Patterns of 1 .. 2,600 bytes (every lanes x words-per-lane class of the kernel, lengths at the word and warp
    boundaries), texts of 0 .. 1,500 bytes, thresholds from 0 to unbounded, N / lower-case bytes included."""
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)

    def rnd(n):
        return alpha[rng.integers(0, 4, size=n)].tobytes().decode()

    def mutate(s, rate):
        out = []
        for ch in s:
            u = rng.random()
            if u < rate / 3:
                out.append("ACGT"[rng.integers(0, 4)])
            elif u < 2 * rate / 3:
                out.append(ch + "ACGT"[rng.integers(0, 4)])
            elif u < rate:
                pass
            else:
                out.append(ch)
        return "".join(out)

    lens = [1, 2, 31, 32, 33, 64, 65, 127, 128, 129, 255, 256, 257, 500, 511, 512, 513, 700, 1023, 1024, 1025, 1500, 2047, 2048, 2049,
            2600]
    patterns = [rnd(n) for n in lens]
    patterns[5] = patterns[5][:10] + "N" + patterns[5][11:]
    patterns[13] = patterns[13][:100] + "acgtn" + patterns[13][105:]
    texts = [""]
    for p in patterns:
        texts.append(rnd(int(rng.integers(0, 60))) + mutate(p, float(rng.choice([0.0, 0.05, 0.15]))) + rnd(int(rng.integers(0, 60))))
    texts += [rnd(int(n)) for n in rng.integers(1, 1500, size=6)]
    texts[3] = texts[3][:5] + "N" + texts[3][6:]
    ks = [int(rng.choice([0, 1, len(p) // 10, len(p) // 5, -1])) for p in patterns]
    got = specimine.hw_distances(patterns, ks, texts)
    want = _oracle_distances(patterns, ks, texts)
    bad = np.argwhere(got != want)
    assert bad.size == 0, [(int(i), int(j), len(patterns[i]), len(texts[j]), ks[i], int(got[i, j]), int(want[i, j])) for i, j in bad[:10]]

"""Native ingest / split / columnar file (csrc/ingest.cpp) against the Python restatement of the same
reference rules (carskit_b200/data.py), which is itself pinned on the reference's sample files (tests/test_data.py)."""
import os

import numpy as np
import pytest

from carskit_b200 import capi, data, synth

REF = "/root/reference"
SAMPLE = os.path.join(REF, "sampleData", "train_binary.csv")


def same_training_set(a: capi.TrainingSet, b: capi.TrainingSet):
    assert (a.num_users, a.num_items, a.num_conditions, a.num_contexts) == (b.num_users, b.num_items, b.num_conditions, b.num_contexts)
    for name in ("u", "j", "ctx", "r", "ctx_ptr", "ctx_cond"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.global_mean == b.global_mean
    assert tuple(a.rating_scale) == tuple(b.rating_scale) and a.num_context_dims == b.num_context_dims
    assert np.array_equal(a.pair_ids, b.pair_ids)


def write_binary_csv(path, ts, rng, duplicates=True):
    """A binary-format file in SHUFFLED line order with a few duplicate (user, item, context) lines and a zero rating."""
    C = ts.num_conditions
    lines = []
    for n in rng.permutation(ts.nnz):
        flags = np.zeros(C, dtype=int)
        flags[ts.ctx_cond[ts.ctx_ptr[ts.ctx[n]]:ts.ctx_ptr[ts.ctx[n] + 1]]] = 1
        lines.append(f"u{ts.u[n]},i{ts.j[n]},{ts.r[n]:g}," + ",".join(map(str, flags)))
    if duplicates:
        lines += [lines[3].rsplit(",", C)[0].rsplit(",", 1)[0] + ",2.5," + lines[3].split(",", 3)[3], lines[7]]
        lines.append(lines[11].rsplit(",", C)[0].rsplit(",", 1)[0] + ",0," + lines[11].split(",", 3)[3])  # zero: dropped
    hdr = "user,item,rating," + ",".join(f"d{c % 3}:c{c}" if c % 4 else f"d{c % 3}:na" for c in range(C))
    open(path, "w").write(hdr + "\n" + "\n".join(lines) + "\n")


@pytest.mark.skipif(not os.path.exists(SAMPLE), reason="reference sample data not on this box")
def test_native_reader_matches_the_reference_sample_layout(cars_lib):
    ds = capi.Dataset.read_binary_csv(SAMPLE)
    ts = ds.training_set()
    want, dao = data.read_binary_csv(SAMPLE)
    same_training_set(ts, want)
    v = ds.view()
    assert (v.num_pairs, v.num_empty_conditions) == (dao.numUserItems(), len(dao.EmptyContextConditions))
    # the hand trace of SURVEY.md section 4
    assert (ts.num_users, ts.num_items, ts.num_contexts, ts.num_conditions, ts.nnz) == (17, 2, 8, 10, 20) and ts.global_mean == 3.95


def test_native_reader_splitter_and_columnar_file_match_the_python_restatement(cars_lib, tmp_path):
    rng = np.random.default_rng(5)
    base, _ = synth.make_training_set(60, 40, [3, 4, 2], 1500, seed=9, order="shuffled")
    path = str(tmp_path / "ratings.csv")
    write_binary_csv(path, base, rng)
    ds = capi.Dataset.read_binary_csv(path)
    ts = ds.training_set()
    want, dao = data.read_binary_csv(path)
    same_training_set(ts, want)
    # columnar round trip
    col = str(tmp_path / "ratings.carscol")
    ds.save(col)
    same_training_set(capi.Dataset.load(col).training_set(), want)
    # DataSplitter: every fold, same labels as the Python restatement (java.util.Random(seed) draws + argsort)
    sp = data.DataSplitter(want, 5, 1)
    for k in range(1, 6):
        tr, te = ds.kfold(5, 1, k)
        wtr, wte = sp.getKthFold(k)
        a, b = tr.training_set(), te.training_set()
        for name in ("u", "j", "ctx", "r"):
            assert np.array_equal(getattr(a, name), getattr(wtr, name)), (k, name)
            assert np.array_equal(getattr(b, name), wte[name]), (k, name)
        assert a.global_mean == wtr.global_mean
    with pytest.raises(capi.CarsError):
        ds.kfold(5, 1, 6)
    with pytest.raises(capi.CarsError):
        capi.Dataset.load(path)  # a CSV is not a columnar file
    with pytest.raises(capi.CarsError):
        capi.Dataset.read_binary_csv(str(tmp_path / "missing.csv"))


def test_from_arrays_round_trip(cars_lib, tmp_path):
    ts, _ = synth.make_training_set(500, 80, [4, 4], 20000, seed=2)
    ds = capi.Dataset.from_training_set(ts, num_context_dims=2)
    col = str(tmp_path / "synthetic.carscol")
    ds.save(col)
    back = capi.Dataset.load(col).training_set()
    for name in ("u", "j", "ctx", "r", "ctx_ptr", "ctx_cond"):
        assert np.array_equal(getattr(back, name), getattr(ts, name)), name
    assert back.global_mean == ts.global_mean and back.rating_scale == (1.0, 5.0)

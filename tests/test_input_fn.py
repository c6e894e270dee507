"""CPU: the host input pipeline mirror (trainers/ml_100k.py:42-61).  get_gpu_input_fn selects records by index with
`_train_index_stream`; it must emit rows in exactly the order get_input_fn's shuffle(16*B).repeat().batch(B) does, so
that the GPU-decoded and the host-parsed pipelines see the same batches for the same seed."""
import numpy as np

from recommender_tensorflow_b200.trainers import ml_100k


def _rows_via_input_fn(path, batch_size, seed, n_batches):
    it = ml_100k.get_input_fn(path, ml_100k.ModeKeys.TRAIN, batch_size=batch_size, seed=seed)()
    out = []
    for _ in range(n_batches):
        feats, _ = next(it)
        out.extend(zip(feats["user_id"].tolist(), feats["item_id"].tolist(), feats["timestamp"].tolist()))
    return out


def test_index_stream_matches_input_fn_order(tmp_path):
    path = str(tmp_path / "t.csv")
    n_rows, batch = 500, 8            # shuffle buffer 128 < rows: replacement phase, drain at the epoch end, repeat
    ml_100k.write_synthetic_csv(path, n_rows)
    rows = list(ml_100k._parse_rows(path))
    key = lambda r: (int(r[0]), int(r[1]), int(r[3]))
    want = _rows_via_input_fn(path, batch, seed=3, n_batches=150)          # 1 200 rows: more than two epochs
    stream = ml_100k._train_index_stream(n_rows, 16 * batch, np.random.default_rng(3))
    got = [key(rows[next(stream)]) for _ in range(len(want))]
    assert got == want


def test_eval_input_fn_keeps_order_and_last_partial_batch(tmp_path):
    path = str(tmp_path / "e.csv")
    ml_100k.write_synthetic_csv(path, 21)
    batches = list(ml_100k.get_input_fn(path, ml_100k.ModeKeys.EVAL, batch_size=8)())
    assert [len(y) for _, y in batches] == [8, 8, 5]
    rows = list(ml_100k._parse_rows(path))
    assert np.concatenate([f["user_id"] for f, _ in batches]).tolist() == [int(r[0]) for r in rows]


def test_index_stream_empty_file_terminates():
    assert list(ml_100k._train_index_stream(0, 16, np.random.default_rng(0))) == []

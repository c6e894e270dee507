"""GPU: the CSV record decoder (dfm_csv_*) against Python's csv module, the restatement of tf.decode_csv the host
input_fn uses (trainers/ml_100k.py:44-58): every consumed column bit-exact, including quoting, defaults, CRLF and
error behaviour; and the decoded device batch drives the train step to the same bits as the host-parsed one."""
import csv
import io

import numpy as np
import pytest

from recommender_tensorflow_b200.csv_reader import GpuCsvReader
from recommender_tensorflow_b200.engine import DeepFMEngine
from recommender_tensorflow_b200.trainers import ml_100k

pytestmark = pytest.mark.gpu


def _engine(**kw):
    fc = ml_100k.get_feature_columns(4)
    return DeepFMEngine(fc["linear"], (), embedding_size=4, hidden_units=(16, 16), feature_dtypes=ml_100k.FEATURE_DTYPES, **kw)


def _host_parse(text, cutoff=5):
    rows = list(csv.reader(io.StringIO(text, newline="")))
    return ml_100k._to_batch(rows, cutoff)


def _assert_same(reader, feats, labels, eng):
    for s in eng.specs:
        name = s["source"]
        got, want = reader.column(name), feats[name]
        if want.dtype == object:
            assert list(got) == list(want), name
        else:
            assert got.dtype == np.int32 and np.array_equal(got, want), name
    assert np.array_equal(reader.labels(), labels)


def _rows(n, rng):
    """records with awkward content in both consumed and skipped fields"""
    out = []
    occ = ["engineer", "", "writer", 'he said ""hi""', "a,b", "none"]
    for i in range(n):
        row = []
        for name, d in zip(ml_100k.COLUMNS, ml_100k.DEFAULTS):
            if name in ml_100k.GENRE:
                row.append(str(int(rng.integers(0, 2))))
            elif name == "title":
                row.append(['"Movie, The (1995)"', '"Say ""Cheese"""', "Plain", ""][i % 4])
            elif name == "occupation":
                v = occ[i % len(occ)]
                row.append('"%s"' % v if ("," in v or '"' in v) else v)
            elif name == "gender":
                row.append(["F", "M", "", '"M"'][i % 4])
            elif name == "zipcode":
                row.append(["55414", "", "V5A2B", "00000"][i % 4])
            elif name == "rating":
                row.append(str(1 + i % 5))
            elif isinstance(d[0], int):
                row.append(["%d" % rng.integers(0, 3000), "", ' 7 ', '"42"', "-1"][i % 5] if name in ("user_id", "item_id", "age", "release_year")
                           else str(int(rng.integers(0, 100))))
            else:
                row.append(["null", "", "x y"][i % 3])
        out.append(",".join(row))
    return out


@pytest.mark.parametrize("n,eol,trailing", [(1, "\n", True), (37, "\n", True), (1000, "\r\n", True), (513, "\n", False)])
def test_csv_decode_equals_python_csv(n, eol, trailing):
    eng = _engine(max_batch=1024)
    rd = GpuCsvReader(eng, ml_100k.COLUMNS, ml_100k.DEFAULTS, ml_100k.LABEL_COL, cutoff=5, max_records=1024)
    rows = _rows(n, np.random.default_rng(n))
    text = eol.join(rows) + (eol if trailing else "")
    pb = rd.decode(text.encode())
    assert pb.batch_size == n
    feats, labels = _host_parse(text)
    _assert_same(rd, feats, labels, eng)
    # the decoded batch feeds K1 directly: same ids as the host-parsed columns
    assert np.array_equal(eng.transform(pb), eng.transform(feats))


def test_csv_decode_synthetic_file_and_train_step(tmp_path):
    path = str(tmp_path / "train.csv")
    ml_100k.write_synthetic_csv(path, 3000)
    a, b = _engine(max_batch=256), _engine(max_batch=256)
    a.init_random(11); b.init_random(11)
    host = ml_100k.get_input_fn(path, ml_100k.ModeKeys.TRAIN, batch_size=256, seed=5)()
    gpu = ml_100k.get_gpu_input_fn(path, b, ml_100k.ModeKeys.TRAIN, batch_size=256, seed=5)()
    for step in range(14):             # crosses the epoch boundary (3000 rows): shuffle-buffer drain included
        feats, y = next(host)
        pb, _ = next(gpu)
        la, lga = a.train_step(feats, y, return_logits=True)
        lb, lgb = b.train_step_device(pb, return_logits=True)
        assert la == lb and np.array_equal(lga, lgb), step
    a.flush(); b.flush()
    for v in a.variable_names():
        assert np.array_equal(a.get_tensor(v), b.get_tensor(v)), v


def test_csv_eval_mode_batches_and_empty_input():
    eng = _engine(max_batch=64)
    rd = GpuCsvReader(eng, ml_100k.COLUMNS, ml_100k.DEFAULTS, ml_100k.LABEL_COL, max_records=64)
    assert rd.decode(b"").batch_size == 0
    rows = _rows(64, np.random.default_rng(3))
    pb = rd.decode(("\n".join(rows) + "\n").encode())
    assert pb.batch_size == 64
    with pytest.raises(Exception):
        rd.decode(("\n".join(_rows(65, np.random.default_rng(4))) + "\n").encode())      # more records than max_records


@pytest.mark.parametrize("bad,what", [
    (lambda r: r + ",1", "fields"),                                 # 43 fields
    (lambda r: ",".join(r.split(",")[:-1]), "fields"),             # 41 fields
    (lambda r: "12x" + r[r.index(","):], "int32"),                 # user_id not an integer
    (lambda r: "99999999999" + r[r.index(","):], "int32"),         # overflow
    (lambda r: r.replace("Plain", 'Pl"ain'), "quot"),              # quote inside an unquoted field
    (lambda r: r.replace("Plain", '"Plain'), "quot|fields"),       # unterminated quote
])
def test_csv_errors_like_decode_csv(bad, what):
    import re
    eng = _engine(max_batch=16)
    rd = GpuCsvReader(eng, ml_100k.COLUMNS, ml_100k.DEFAULTS, ml_100k.LABEL_COL, max_records=16)
    rows = _rows(8, np.random.default_rng(9))
    rows[2] = "5,7,3,0,null,0,0,0,0,0,30,M,writer,55414,null,null,null,Plain,null,null,null," + ",".join(["0"] * 19) + ",null,1990"
    assert rd.decode(("\n".join(rows) + "\n").encode()).batch_size == 8
    rows[2] = bad(rows[2])
    with pytest.raises(ValueError) as ei:
        rd.decode(("\n".join(rows) + "\n").encode())
    assert "record 2" in str(ei.value) and re.search(what, str(ei.value)), str(ei.value)


def test_csv_decode_full_batch_properties():
    """65 536 records: counts, label rate and id histograms equal the host parse of the same text."""
    eng = _engine(max_batch=65536)
    rd = GpuCsvReader(eng, ml_100k.COLUMNS, ml_100k.DEFAULTS, ml_100k.LABEL_COL, max_records=65536)
    rows = _rows(4096, np.random.default_rng(1)) * 16
    text = "\n".join(rows) + "\n"
    pb = rd.decode(text.encode())
    assert pb.batch_size == 65536
    feats, labels = _host_parse(text)
    _assert_same(rd, feats, labels, eng)


def test_estimator_with_gpu_input_equals_host_input(tmp_path):
    """trainers.deep_fm.train_and_evaluate --gpu-input: same record stream, same bits as the host-parsed run."""
    from types import SimpleNamespace
    from recommender_tensorflow_b200.trainers import deep_fm
    train, test = str(tmp_path / "train.csv"), str(tmp_path / "test.csv")
    ml_100k.write_synthetic_csv(train, 1500, seed=1)
    ml_100k.write_synthetic_csv(test, 400, seed=2)
    out = []
    for gpu in (False, True):
        args = SimpleNamespace(train_csv=train, test_csv=test, job_dir=str(tmp_path / ("job%d" % gpu)), restore=False, embedding_size=4,
                               hidden_units=[16, 16], dropout=0, batch_size=64, train_steps=30, exclude_linear=False, exclude_mf=False,
                               exclude_dnn=False, gpu_input=gpu, seed=7)
        np.random.seed(0)
        out.append(deep_fm.train_and_evaluate(args))
    assert out[0]["global_step"] == out[1]["global_step"] == 30
    for k in ("average_loss", "accuracy", "auc"):
        assert out[0][k] == out[1][k], (k, out[0][k], out[1][k])


def test_csv_file_resident_decode_lines():
    """dfm_csv_load + dfm_csv_decode_lines: the file is split into lines on the device once; a batch is a list of line
    numbers (with repeats, in any order) and decodes to the same columns as the host parse of those lines."""
    eng = _engine(max_batch=512)
    rows = _rows(300, np.random.default_rng(21))
    header = ",".join(ml_100k.COLUMNS)
    for eol, trailing in (("\n", True), ("\r\n", False)):
        text = eol.join([header] + rows) + (eol if trailing else "")
        rd = GpuCsvReader(eng, ml_100k.COLUMNS, ml_100k.DEFAULTS, ml_100k.LABEL_COL, max_records=512, max_bytes=512 * 400)
        assert rd.load_file(text.encode()) == 301
        idx = np.random.default_rng(22).integers(0, 300, 512)
        pb = rd.decode_lines(idx + 1)
        assert pb.batch_size == 512
        feats, labels = _host_parse("\n".join(rows[i] for i in idx) + "\n")
        _assert_same(rd, feats, labels, eng)
        with pytest.raises(Exception):
            rd.decode_lines(np.array([301]))          # past the last line
        with pytest.raises(ValueError):
            rd.decode_lines(np.array([0]))            # the header is not a record: "user_id" is not an int32
        rd.close()

"""debug: configs[2] at B=65536, compare every tensor after N steps with the oracle pair; env switches bisect the path"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_tensorflow_b200 import synth
from recommender_tensorflow_b200.engine import DeepFMEngine
from tests.util import make_pair, ml100k_columns

B = int(os.environ.get("DBG_B", "65536")); steps = int(os.environ.get("DBG_STEPS", "1"))
hidden = tuple(int(x) for x in os.environ.get("DBG_H", "256,128").split(","))
cols, dtypes = ml100k_columns()
eng = DeepFMEngine(cols, (), embedding_size=16, hidden_units=hidden, max_batch=B, feature_dtypes=dtypes)
ora, _ = make_pair(eng, seed=21)
ml, rng = synth.ML100K(), np.random.default_rng(22)
for i in range(steps):
    f, y = ml.batch(B, rng)
    loss, logits = eng.train_step(f, y, return_logits=True)
    rloss, rlogits = ora.train_step_raw(f, y)
    l64, z64 = ora.last64
    print("step", i, "loss", loss, rloss, l64, "logits err", np.abs(logits - z64).max(), "oracle32 err", np.abs(rlogits - z64).max())
st, r32, r64 = eng.state(), ora.state(), ora.state64()
offs = eng.row_offsets
names = [s["name"] for s in eng.specs]
for name in r32:
    got = st[name].reshape(r32[name].shape).astype(np.float64)
    e = np.abs(got - r64[name]); e32 = np.abs(r32[name].astype(np.float64) - r64[name])
    print("%-8s max err %.3e  oracle32 err %.3e  scale %.3e" % (name, e.max(), e32.max(), np.abs(r64[name]).max()))
    if name in ("emb", "emb/m", "emb/v", "lin"):
        rows = e.reshape(e.shape[0], -1).max(1)
        worst = np.argsort(-rows)[:6]
        for r in worst:
            f_ = int(np.searchsorted(offs, r, side="right") - 1)
            print("     row %d (field %s, id %d): err %.3e, oracle32 err %.3e" % (r, names[f_], r - offs[f_], rows[r], e32.reshape(e32.shape[0], -1).max(1)[r]))

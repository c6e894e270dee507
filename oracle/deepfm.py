"""ORACLE (test infrastructure): CPU restatement of the reference train step.  PARITY UNPINNED for
floating point (no TF-1.12 available, the reference has no tests) — see oracle/__init__.py.

Follows, line by line:
  trainers/deep_fm.py:37-45    linear block     z_lin = sum_f w_f[id_f] + sum_j x_j wn_j + bias
  trainers/deep_fm.py:47-74    input layer      E[b,f,:] = emb_f[id_f];  E[b,dc+j,:] = x_j * Vnum[j,:]
  trainers/deep_fm.py:76-91    FM block         z_fm = 0.5 * sum_k((sum_f E)^2 - sum_f E^2)
  trainers/deep_fm.py:93-112   DNN block        relu dense stack + dense(1)
  trainers/deep_fm.py:114-125  head             sigmoid CE, mean over batch (contrib binary head)
  trainers/model_utils.py:57-66 optimizer       tf.train.AdamOptimizer(lr) -> NON-LAZY sparse Adam on
                                                 tables, ApplyAdam on dense vars
  trainers/linear_deep.py:32-39 wide&deep       no FM, SUM loss, Adagrad (dnn side) + FTRL (linear side)

All arithmetic is float32 (dtype=float64 available for tolerance calibration).  Initial weights
are injected by the caller: TF's initializer RNG is not reproducible outside TF.

Generic variable names (shared with the product's C ABI; the host facade maps TF-1.12 names):
  emb [R,k]  lin [R]  (R = sum of per-field bucket counts, fields concatenated in model order)
  num_emb [dn,k]  num_lin [dn]  bias [1]  W0..W{L-1}, b0..b{L-1}  Wo [hL,1]  bo [1]
"""
import numpy as np
import torch

from .transforms import num_buckets, transform


def default_opt(name="Adam", lr=0.001):
    d = dict(name=name, lr=lr)
    if name == "Adam":
        d.update(beta1=0.9, beta2=0.999, eps=1e-8)
    elif name == "Adagrad":
        d.update(init_acc=0.1)
    elif name == "Ftrl":
        d.update(init_acc=0.1, lr_power=-0.5, l1=0.0, l2=0.0)
    elif name == "RMSProp":   # tf.train.RMSPropOptimizer defaults (python/training/rmsprop.py)
        d.update(beta1=0.0, beta2=0.9, eps=1e-10, init_acc=1.0)
    return d


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (the product's dropout generator, mlp_kernels.cuh::dfm_mix64)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def dropout_mask(seed, step, layer, rows, width, keep):
    """keep ? 1/keep : 0 for every element (row-major index) of a [rows, width] activation.  Restates
    tf.layers.dropout(net, rate, training=True) (trainers/deep_fm.py:102-103) with the product's own
    counter-based generator - TF's stream is not reproducible."""
    with np.errstate(over="ignore"):
        key = _mix64(_mix64(np.uint64(seed)) ^ np.uint64(step * 64 + layer))
        idx = np.arange(rows * width, dtype=np.uint64)
        u = (_mix64(key ^ idx) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.where(u < np.float32(keep), np.float32(1.0 / np.float32(keep)), np.float32(0)).reshape(rows, width)


class Var:
    def __init__(self, value, group):
        self.w = value
        self.group = group  # "deep" | "linear" -> which optimizer
        self.slots = {}


class OracleDeepFM:
    def __init__(self, cfg, weights, dtype=torch.float32):
        """cfg keys: cat (specs, model order), num (names), k, hidden, use_linear/use_mf/use_dnn,
        loss_reduction ('mean'|'sum'), opt_deep, opt_linear.  weights: dict generic-name -> ndarray."""
        self.cfg = cfg
        self.dt = dtype
        self.cat = cfg["cat"]
        self.num = cfg.get("num", [])
        self.k = cfg["k"]
        self.hidden = list(cfg.get("hidden", [16, 16]))
        self.use_linear = bool(cfg.get("use_linear", True))
        self.use_mf = bool(cfg.get("use_mf", True))
        self.use_dnn = bool(cfg.get("use_dnn", True))
        self.red = cfg.get("loss_reduction", "mean")
        self.dropout = float(cfg.get("dropout", 0.0))
        self.dropout_seed = int(cfg.get("dropout_seed", 0))
        # params["activation"] of model_fn (trainers/deep_fm.py:22, handed to tf.layers.dense at :100; default tf.nn.relu)
        self.activation = cfg.get("activation", "relu") or "identity"
        self.opt = {"deep": cfg.get("opt_deep", default_opt()),
                    "linear": cfg.get("opt_linear", default_opt())}
        self.nb = [num_buckets(s) for s in self.cat]
        self.off = np.concatenate([[0], np.cumsum(self.nb)]).astype(np.int64)
        self.R = int(self.off[-1])
        self.dc, self.dn = len(self.cat), len(self.num)
        # multivalent columns: slot -> field; embedding of a field = mean of its present slots, linear = sum
        self.width = [int(s.get("width", 1)) for s in self.cat]
        self.slot_field = np.repeat(np.arange(self.dc), self.width)
        self.n_slots = int(self.slot_field.shape[0])
        self.P = torch.zeros(self.n_slots, self.dc, dtype=dtype)     # slot -> field pooling matrix
        self.P[torch.arange(self.n_slots), torch.as_tensor(self.slot_field)] = 1
        self.vars = {}
        group = {"lin": "linear", "num_lin": "linear", "bias": "linear"}
        for name, val in weights.items():
            self.vars[name] = Var(torch.tensor(np.asarray(val), dtype=dtype).clone(),
                                  group.get(name, "deep"))
        self.t = 0
        f32 = np.float32
        self.pow = {g: [f32(1.0), f32(1.0)] for g in ("deep", "linear")}  # beta1^t, beta2^t (fp32 products)
        for v in self.vars.values():
            o = self.opt[v.group]
            if o["name"] == "Adam":
                v.slots = {"m": torch.zeros_like(v.w), "v": torch.zeros_like(v.w)}
            elif o["name"] == "Adagrad":
                v.slots = {"acc": torch.full_like(v.w, o["init_acc"])}
            elif o["name"] == "Ftrl":
                v.slots = {"acc": torch.full_like(v.w, o["init_acc"]), "lin": torch.zeros_like(v.w)}
            elif o["name"] == "RMSProp":
                v.slots = {"rms": torch.full_like(v.w, o["init_acc"]), "mom": torch.zeros_like(v.w)}
            elif o["name"] == "SGD":
                v.slots = {}
            else:
                raise ValueError(o["name"])

    # ---------------------------------------------------------------- forward
    def rows(self, ids):
        """ids [B,dc] int (-1 empty) -> global row index [B,dc] (clamped) and validity mask."""
        ids = torch.as_tensor(np.asarray(ids), dtype=torch.int64)
        valid = ids >= 0
        rows = ids.clamp(min=0) + torch.as_tensor(self.off[:-1][self.slot_field])[None, :]
        return rows, valid

    def forward(self, ids, x=None, keep=False, train=False):
        W = {n: v.w for n, v in self.vars.items()}
        rows, valid = self.rows(ids)
        B = rows.shape[0]
        dt = self.dt
        vm = valid.to(dt)
        x = torch.zeros(B, 0, dtype=dt) if x is None else torch.as_tensor(np.asarray(x), dtype=dt)
        z = torch.zeros(B, dtype=dt)
        cache = dict(rows=rows, valid=valid, x=x)
        cnt = vm @ self.P                                     # [B, dc] present slots per field
        inv = torch.where(cnt > 0, 1.0 / cnt.clamp(min=1), torch.zeros_like(cnt))
        cache["inv"] = inv
        if self.use_linear:
            z_lin = (W["lin"][rows] * vm).sum(1)
            if self.dn:
                z_lin = z_lin + x @ W["num_lin"]
            z = z + (z_lin + W["bias"][0])
        if self.use_mf or self.use_dnn:
            Es = W["emb"][rows] * vm[:, :, None]                      # [B,n_slots,k]
            E = torch.einsum("bsk,sf->bfk", Es, self.P) * inv[:, :, None]   # mean over the present slots -> [B,dc,k]
            if self.dn:
                E = torch.cat([E, x[:, :, None] * W["num_emb"][None]], 1)   # [B,d,k]
            cache["E"] = E
        if self.use_mf:
            s = E.sum(1)
            z = z + 0.5 * (s * s - (E * E).sum(1)).sum(1)
            cache["s"] = s
        if self.use_dnn:
            h = E.reshape(B, -1)
            acts, pre, masks = [h], [], []
            fwd = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "identity": lambda v: v}[self.activation]
            for i in range(len(self.hidden)):
                h = fwd(h @ W["W%d" % i] + W["b%d" % i])
                pre.append(h)                                   # activation output before dropout
                mk = None
                if train and self.dropout > 0:
                    keepp = np.float32(1.0) - np.float32(self.dropout)
                    mk = torch.as_tensor(dropout_mask(self.dropout_seed, self.t + 1, i, B, h.shape[1], keepp), dtype=dt)
                    h = h * mk
                masks.append(mk)
                acts.append(h)
            cache["pre"], cache["masks"] = pre, masks
            z = z + (h @ W["Wo"])[:, 0] + W["bo"][0]
            cache["acts"] = acts
        return (z, cache) if keep else z

    def layer_tensors(self, ids, x=None, train=True):
        """The tensors the reference hands to layer_summary (trainers/deep_fm.py:43,89,105,110,115), by TF name scope."""
        W = {n: v.w for n, v in self.vars.items()}
        z, c = self.forward(ids, x, keep=True, train=train)
        out = {}
        if self.use_linear:
            vm = c["valid"].to(self.dt)
            z_lin = (W["lin"][c["rows"]] * vm).sum(1)
            if self.dn:
                z_lin = z_lin + c["x"] @ W["num_lin"]
            out["linear/linear"] = z_lin + W["bias"][0]
        if self.use_mf:
            s, E = c["s"], c["E"]
            out["mf/logits"] = 0.5 * (s * s - (E * E).sum(1)).sum(1)
        if self.use_dnn:
            for i in range(len(self.hidden)):
                out["dnn/dnn/hiddenlayer_%d" % i] = c["acts"][i + 1]
            out["dnn/dnn/logits"] = (c["acts"][-1] @ W["Wo"])[:, 0] + W["bo"][0]
        out["deep_fm/logits"] = z
        return {k: v.detach().numpy() for k, v in out.items()}

    @staticmethod
    def loss_vec(z, y):
        return torch.clamp(z, min=0) - z * y + torch.log1p(torch.exp(-torch.abs(z)))

    # ---------------------------------------------------------------- backward
    def grads(self, ids, x, y):
        z, c = self.forward(ids, x, keep=True, train=True)
        y = torch.as_tensor(np.asarray(y), dtype=self.dt)
        B = z.shape[0]
        lv = self.loss_vec(z, y)
        loss = lv.mean() if self.red == "mean" else lv.sum()
        dz = torch.sigmoid(z) - y
        if self.red == "mean":
            dz = dz / B
        W = {n: v.w for n, v in self.vars.items()}
        g = {}
        rows, valid, x = c["rows"], c["valid"], c["x"]
        flat_rows = rows[valid]
        if self.use_linear:
            g["lin"] = ("sparse", flat_rows, dz[:, None].expand(B, self.n_slots)[valid])
            g["bias"] = dz.sum().reshape(1)
            if self.dn:
                g["num_lin"] = x.t() @ dz
        if self.use_mf or self.use_dnn:
            E = c["E"]
            dE = torch.zeros_like(E)
            if self.use_mf:
                dE = dE + dz[:, None, None] * (c["s"][:, None, :] - E)
            if self.use_dnn:
                acts = c["acts"]
                L = len(self.hidden)
                dh = dz[:, None] * W["Wo"][:, 0][None, :]
                g["Wo"] = acts[L].t() @ dz[:, None]
                g["bo"] = dz.sum().reshape(1)
                fprime = {"relu": lambda yv: (yv > 0).to(self.dt), "tanh": lambda yv: 1 - yv * yv,
                          "sigmoid": lambda yv: yv * (1 - yv), "identity": lambda yv: torch.ones_like(yv)}[self.activation]
                for i in reversed(range(L)):
                    # derivative of the activation in terms of its output, times the dropout factor (0 or 1/keep)
                    dh = dh * fprime(c["pre"][i])
                    if c["masks"][i] is not None:
                        dh = dh * c["masks"][i]
                    g["W%d" % i] = acts[i].t() @ dh
                    g["b%d" % i] = dh.sum(0)
                    dh = dh @ W["W%d" % i].t()
                dE = dE + dh.reshape(E.shape)
            dEs = (dE[:, :self.dc] * c["inv"][:, :, None])[:, self.slot_field]      # every present slot gets dE_field / count
            g["emb"] = ("sparse", flat_rows, dEs[valid])
            if self.dn:
                g["num_emb"] = (x[:, :, None] * dE[:, self.dc:]).sum(0)
        return loss, z, g

    # ---------------------------------------------------------------- optimizers
    def _alpha(self, grp):
        o = self.opt[grp]
        f32 = np.float32
        b1p, b2p = self.pow[grp]
        if self.dt == torch.float64:
            return o["lr"] * np.sqrt(1.0 - float(b2p)) / (1.0 - float(b1p))
        return f32(f32(o["lr"]) * np.sqrt(f32(1) - b2p) / (f32(1) - b1p))

    def _apply(self, var, grad):
        o = self.opt[var.group]
        name = o["name"]
        sparse = isinstance(grad, tuple)
        if sparse:
            _, rows, vals = grad
            uniq, inv = torch.unique(rows, return_inverse=True)
            gsum = torch.zeros((uniq.shape[0],) + tuple(var.w.shape[1:]), dtype=self.dt)
            gsum.index_add_(0, inv, vals)                      # dedup (unique + segment_sum)
        if name == "Adam":
            b1, b2, eps = o["beta1"], o["beta2"], o["eps"]
            a = float(self._alpha(var.group))
            if self.dt == torch.float32:    # TF evaluates (1 - beta) in float32
                omb1 = float(np.float32(1) - np.float32(b1))
                omb2 = float(np.float32(1) - np.float32(b2))
            else:
                omb1, omb2 = 1 - b1, 1 - b2
            m, v = var.slots["m"], var.slots["v"]
            if sparse:                                          # NON-LAZY: whole variable moves
                m.mul_(b1); m[uniq] += gsum * omb1
                v.mul_(b2); v[uniq] += (gsum * gsum) * omb2
                var.w -= (a * m) / (v.sqrt() + eps)
            else:                                               # ApplyAdam
                m += (grad - m) * omb1
                v += (grad * grad - v) * omb2
                var.w -= (m * a) / (v.sqrt() + eps)
        elif name == "Adagrad":
            acc = var.slots["acc"]
            if sparse:
                acc[uniq] += gsum * gsum
                var.w[uniq] -= o["lr"] * gsum / acc[uniq].sqrt()
            else:
                acc += grad * grad
                var.w -= o["lr"] * grad / acc.sqrt()
        elif name == "Ftrl":
            acc, lin = var.slots["acc"], var.slots["lin"]
            lr = o["lr"]
            idx = uniq if sparse else slice(None)
            gg = gsum if sparse else grad
            new_acc = acc[idx] + gg * gg
            lin[idx] = lin[idx] + gg - (new_acc.sqrt() - acc[idx].sqrt()) / lr * var.w[idx]
            var.w[idx] = -lin[idx] / (new_acc.sqrt() / lr)
            acc[idx] = new_acc
        elif name == "RMSProp":   # core/kernels/training_ops.cc ApplyRMSProp / SparseApplyRMSProp (touched rows only)
            rms, mom = var.slots["rms"], var.slots["mom"]
            idx = uniq if sparse else slice(None)
            gg = gsum if sparse else grad
            omr = float(np.float32(1) - np.float32(o["beta2"])) if self.dt == torch.float32 else 1 - o["beta2"]
            rms[idx] = rms[idx] + (gg * gg - rms[idx]) * omr
            mom[idx] = mom[idx] * o["beta1"] + (gg * o["lr"]) / (rms[idx] + o["eps"]).sqrt()
            var.w[idx] = var.w[idx] - mom[idx]
        elif name == "SGD":
            if sparse:
                var.w[uniq] -= o["lr"] * gsum
            else:
                var.w -= o["lr"] * grad
        else:
            raise ValueError(name)

    def train_step(self, ids, x, y):
        """One reference train step.  Returns (loss, logits) evaluated BEFORE the update."""
        loss, z, g = self.grads(ids, x, y)
        self.t += 1
        f32 = np.float32
        for grp in ("deep", "linear"):
            o = self.opt[grp]
            if o["name"] == "Adam":   # powers as seen by step t: beta^t (TF updates them after apply)
                self.pow[grp] = [f32(self.pow[grp][0] * f32(o["beta1"])),
                                 f32(self.pow[grp][1] * f32(o["beta2"]))]
        for name, grad in g.items():
            self._apply(self.vars[name], grad)
        return float(loss), z.detach().numpy().copy()

    def train_step_raw(self, features, labels):
        ids = transform(self.cat, features)
        x = None
        if self.dn:
            x = np.stack([np.asarray(features[n], dtype=np.float32) for n in self.num], 1)
        return self.train_step(ids, x, labels)

    def state(self):
        out = {}
        for n, v in self.vars.items():
            out[n] = v.w.numpy().copy()
            for sn, sv in v.slots.items():
                out[n + "/" + sn] = sv.numpy().copy()
        return out


# ----------------------------------------------------------------------------------------------
def init_weights(cfg, seed=0):
    """Reference-style initial weights: tables truncated-normal(0, 1/sqrt(k)) cut at
    2 sigma, linear zeros, dense kernels + numeric_embeddings glorot-uniform, biases zeros.
    Own generator (numpy) - both oracle and CUDA side load exactly these arrays."""
    rng = np.random.default_rng(seed)
    k = cfg["k"]
    nb = [num_buckets(s) for s in cfg["cat"]]
    R, dc, dn = int(sum(nb)), len(cfg["cat"]), len(cfg.get("num", []))
    w = {}
    if cfg.get("use_linear", True):
        w["lin"] = np.zeros(R, np.float32)
        w["bias"] = np.zeros(1, np.float32)
        if dn:
            w["num_lin"] = np.zeros(dn, np.float32)
    if cfg.get("use_mf", True) or cfg.get("use_dnn", True):
        sigma = 1.0 / np.sqrt(k)
        e = rng.standard_normal((R, k))
        bad = np.abs(e) > 2
        while bad.any():
            e[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(e) > 2
        w["emb"] = (e * sigma).astype(np.float32)
        if dn:
            lim = np.sqrt(6.0 / (dn + k))
            w["num_emb"] = rng.uniform(-lim, lim, (dn, k)).astype(np.float32)
    if cfg.get("use_dnn", True):
        fan_in = (dc + dn) * k
        for i, h in enumerate(cfg.get("hidden", [16, 16])):
            lim = np.sqrt(6.0 / (fan_in + h))
            w["W%d" % i] = rng.uniform(-lim, lim, (fan_in, h)).astype(np.float32)
            w["b%d" % i] = np.zeros(h, np.float32)
            fan_in = h
        lim = np.sqrt(6.0 / (fan_in + 1))
        w["Wo"] = rng.uniform(-lim, lim, (fan_in, 1)).astype(np.float32)
        w["bo"] = np.zeros(1, np.float32)
    return w

"""GPU CSV input path: whole records in, device-resident raw batch out.

`GpuCsvReader` binds `dfm_csv_*` (include/deepfm_b200.h): the equivalent of the reference's
`tf.data.TextLineDataset(csv)` + `tf.decode_csv(value, DEFAULTS)` + `rating >= cutoff`
(trainers/ml_100k.py:44-58) with the field work done by CUDA kernels.  `decode()` returns a
`PackedBatch` whose column pointers alias the reader's device buffers (valid until the next decode),
ready for `DeepFMEngine.train_step_device / transform / predict_logits`.
"""
import ctypes as C

import numpy as np

from . import _lib
from .engine import DfmError, PackedBatch

CSV_SKIP, CSV_INT32, CSV_STRING = 0, 1, 2


class CsvConfig(C.Structure):
    _fields_ = [("n_fields", C.c_int32), ("kind", C.POINTER(C.c_int32)), ("int_default", C.POINTER(C.c_int32)),
                ("str_default", C.POINTER(C.c_char_p)), ("label_field", C.c_int32), ("label_min", C.c_int32),
                ("max_records", C.c_int32), ("max_bytes", C.c_int64), ("device", C.c_int32)]


class GpuCsvReader:
    def __init__(self, engine, columns, defaults, label_col=None, cutoff=5, max_records=65536, max_bytes=None):
        """columns / defaults: the CSV schema (names and `record_defaults`, e.g. trainers.ml_100k.COLUMNS / DEFAULTS).
        Only the fields `engine` consumes (plus the label) are materialised; the rest is parsed for structure only."""
        self.lib = _lib.load()
        self.eng = engine
        self.columns = list(columns)
        if engine.num_columns:
            raise ValueError("GpuCsvReader: numeric (float) model columns are not supported, the CSV schema is int32 / string")
        wanted = {}
        for s in engine.specs:
            if s["width"] != 1:
                raise ValueError("GpuCsvReader: multivalent column %r has no CSV encoding in the reference" % s["source"])
            wanted[s["source"]] = s["dtype"]
        n = len(self.columns)
        kind = (C.c_int32 * n)()
        idef = (C.c_int32 * n)()
        sdef = (C.c_char_p * n)()
        for j, (name, d) in enumerate(zip(self.columns, defaults)):
            is_int = isinstance(d[0], (int, np.integer))
            if name in wanted and wanted[name] != ("int32" if is_int else "string"):
                raise ValueError("column %r: model expects %s, CSV default %r says otherwise" % (name, wanted[name], d))
            if name in wanted or name == label_col:
                kind[j] = CSV_INT32 if is_int else CSV_STRING
            else:
                kind[j] = CSV_SKIP
            if is_int:
                idef[j] = int(d[0])
            else:
                sdef[j] = str(d[0]).encode()
        missing = [k for k in wanted if k not in self.columns]
        if missing:
            raise ValueError("model columns missing from the CSV schema: %r" % missing)
        self.label_field = self.columns.index(label_col) if label_col is not None else -1
        self.max_records = int(max_records)
        self.max_bytes = int(max_bytes or 512 * self.max_records)
        cfg = CsvConfig(n, C.cast(kind, C.POINTER(C.c_int32)), C.cast(idef, C.POINTER(C.c_int32)), C.cast(sdef, C.POINTER(C.c_char_p)),
                        self.label_field, int(cutoff), self.max_records, self.max_bytes, engine.device)
        self._keep = [kind, idef, sdef]
        h = C.c_void_p()
        rc = self.lib.dfm_csv_create(C.byref(cfg), C.byref(h))
        if rc != _lib.DFM_OK:
            raise DfmError(rc, (self.lib.dfm_csv_last_error(None) or b"").decode())
        self.h = h
        self._field = {name: j for j, name in enumerate(self.columns)}

    def close(self):
        if getattr(self, "h", None):
            self.lib.dfm_csv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.DFM_OK:
            msg = (self.lib.dfm_csv_last_error(self.h) or b"").decode()
            if rc == -7:
                raise ValueError(msg)       # tf.decode_csv raises InvalidArgumentError; input errors are ValueErrors here
            raise DfmError(rc, msg)

    def decode(self, text, stream=None):
        """text: bytes / bytearray / numpy uint8 array / pinned torch uint8 tensor holding whole records (host), or a
        cuda uint8 tensor (16-byte aligned, padded to a multiple of 16).  -> PackedBatch on the device."""
        n_rec = C.c_int32(0)
        st = C.c_void_p(stream) if stream else None
        if not stream:
            self.eng.sync()      # the previous batch aliases the buffers this decode overwrites; a step may still be reading them
        if hasattr(text, "is_cuda") and text.is_cuda:
            self._check(self.lib.dfm_csv_decode(self.h, C.c_void_p(text.data_ptr()), int(text.numel()), C.byref(n_rec), st))
        elif hasattr(text, "data_ptr"):
            self._check(self.lib.dfm_csv_decode_host(self.h, C.c_void_p(text.data_ptr()), int(text.numel()), C.byref(n_rec), st))
        else:
            buf = text if isinstance(text, np.ndarray) else np.frombuffer(bytes(text), dtype=np.uint8)
            self._check(self.lib.dfm_csv_decode_host(self.h, C.c_void_p(buf.ctypes.data), int(buf.size), C.byref(n_rec), st))
        return self._batch(n_rec.value)

    def load_file(self, source):
        """File-resident mode: upload the whole CSV (path or bytes) once; -> number of lines (line 0 = header, if any)."""
        data = np.fromfile(source, dtype=np.uint8) if isinstance(source, str) else np.frombuffer(bytes(source), dtype=np.uint8)
        n = C.c_int64(0)
        self._check(self.lib.dfm_csv_load(self.h, C.c_void_p(data.ctypes.data), int(data.size), C.byref(n)))
        return int(n.value)

    def decode_lines(self, line_idx, stream=None):
        """Decode the given lines of the loaded file (record i of the batch = line line_idx[i]) -> PackedBatch."""
        idx = np.ascontiguousarray(line_idx, dtype=np.int32)
        if not stream:
            self.eng.sync()
        self._check(self.lib.dfm_csv_decode_lines(self.h, C.c_void_p(idx.ctypes.data), int(idx.size), C.c_void_p(stream) if stream else None))
        return self._batch(int(idx.size))

    def _batch(self, n):
        eng = self.eng
        nc = max(len(eng.specs), 1)
        cat = (C.c_void_p * nc)()
        off = (C.c_void_p * nc)()
        num = (C.c_void_p * 1)()
        for i, s in enumerate(eng.specs):
            f = self._field[s["source"]]
            if s["dtype"] == "string":
                cat[i] = self.lib.dfm_csv_str_bytes(self.h, f)
                off[i] = self.lib.dfm_csv_str_offsets(self.h, f)
            else:
                cat[i] = self.lib.dfm_csv_int_column(self.h, f)
        lab = self.lib.dfm_csv_labels(self.h) if self.label_field >= 0 else None
        raw = _lib.RawBatch(n, C.cast(cat, C.POINTER(C.c_void_p)), C.cast(off, C.POINTER(C.c_void_p)),
                            C.cast(num, C.POINTER(C.c_void_p)), lab)
        return PackedBatch(None, 0, raw, [cat, off, num, self], 0, n, True)

    # ---- read-back helpers (tests, debugging) ----
    @staticmethod
    def _d2h(ptr, nbytes):
        import torch
        if not nbytes:
            return np.empty(0, dtype=np.uint8)

        class _View:      # device memory owned by the reader, exposed through the CUDA array interface
            __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        torch.cuda.synchronize()
        return torch.as_tensor(_View(), device="cuda").cpu().numpy()

    def column(self, name):
        """Decoded column as numpy: int32 [n] or an object array of bytes (strings)."""
        f = self._field[name]
        n = int(self.lib.dfm_csv_num_records(self.h))
        p = self.lib.dfm_csv_int_column(self.h, f)
        if p:
            return self._d2h(p, 4 * n).view(np.int32).copy()
        offs = self._d2h(self.lib.dfm_csv_str_offsets(self.h, f), 4 * (n + 1)).view(np.int32).copy()
        data = self._d2h(self.lib.dfm_csv_str_bytes(self.h, f), int(offs[-1])).tobytes()
        return np.array([data[offs[i]:offs[i + 1]] for i in range(n)], dtype=object)

    def labels(self):
        n = int(self.lib.dfm_csv_num_records(self.h))
        return self._d2h(self.lib.dfm_csv_labels(self.h), 4 * n).view(np.float32).copy()

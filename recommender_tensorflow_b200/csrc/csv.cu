// GPU CSV record decoder: replaces tf.data.TextLineDataset(csv) + tf.decode_csv(value, DEFAULTS)
// of the reference's input_fn (trainers/ml_100k.py:44-58) for the columns the model consumes.  The host only moves
// whole records (lines) around; field splitting, RFC-4180 unquoting, int32 parsing, default substitution and the
// label threshold run on the device and leave the batch in the layout dfm_raw_batch expects (int32 columns,
// Arrow-style string columns, float labels) — no per-field work on the CPU.
//
// Pipeline of one dfm_csv_decode call (byte / integer work, HBM-bound):
//   csv_count_nl      16 bytes per thread: newline count                      -> exclusive scan (prims)
//   csv_line_starts   same bytes: position after every newline                -> line_start[r]
//   csv_parse         thread per record: walk the fields; ints parsed in place, strings measured (pos, length, flags)
//   exclusive scan    per string column: lengths -> Arrow offsets
//   csv_copy_strings  thread per (record, string column): unquote / default -> bytes
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/deepfm_b200.h"
#include "prims.cuh"

namespace {

constexpr int CSV_CHUNK = 16;      // bytes per thread in the newline kernels

struct CsvDev {
    int n_fields;
    const int32_t* kind;           // [n_fields] DFM_CSV_*
    const int32_t* slot;           // [n_fields] index among the int / string outputs
    const int32_t* int_default;    // [n_fields]
    const int32_t* def_off;        // [n_fields + 1] offsets into def_bytes (string defaults)
    const char* def_bytes;
    int label_field, label_min;
};

__device__ __forceinline__ void csv_fail(unsigned long long* err, uint32_t rec, uint32_t code) {
    atomicMin(err, ((unsigned long long)rec << 8) | code);
}

__global__ void __launch_bounds__(256) csv_count_nl_kernel(const uint4* __restrict__ text16, int64_t n_bytes, uint32_t* __restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = t * CSV_CHUNK;
    if (base >= n_bytes) return;
    uint4 v = text16[t];           // the buffer is padded to a multiple of 16 bytes
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (base + i < n_bytes && ((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) == '\n') ++c;
    counts[t] = c;
}

__global__ void __launch_bounds__(256) csv_line_starts_kernel(const uint4* __restrict__ text16, int64_t n_bytes, const uint32_t* __restrict__ offs,
                                                              uint32_t* __restrict__ line_start, uint32_t max_records) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t base = t * CSV_CHUNK;
    if (base >= n_bytes) return;
    if (t == 0) line_start[0] = 0;
    uint4 v = text16[t];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t k = offs[t];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (base + i < n_bytes && ((w[i >> 2] >> ((i & 3) * 8)) & 0xffu) == '\n') {
            ++k;
            if (k <= max_records) line_start[k] = (uint32_t)(base + i + 1);
        }
}

// One thread per record.  Field grammar = tf.decode_csv (RFC 4180): a field is either unquoted (no '"' inside) or
// quoted with '""' as the escaped quote; an empty field takes the column default; the record must have exactly
// n_fields fields.  int32 fields: optional blanks, sign, digits, optional blanks.
__global__ void __launch_bounds__(128) csv_parse_kernel(const char* __restrict__ text, const uint32_t* __restrict__ line_start,
                                                        const int32_t* __restrict__ line_idx /* null: record r is line r */, uint32_t n_rec,
                                                        CsvDev cfg, int32_t* const* __restrict__ int_out, uint32_t* const* __restrict__ str_pos,
                                                        uint32_t* const* __restrict__ str_len, uint8_t* const* __restrict__ str_flag,
                                                        float* __restrict__ labels, unsigned long long* __restrict__ err) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint32_t li = line_idx ? (uint32_t)line_idx[r] : r;
    uint32_t p = line_start[li], e = line_start[li + 1];
    if (e > p && text[e - 1] == '\n') --e;
    if (e > p && text[e - 1] == '\r') --e;
    int field = 0;
    while (true) {
        uint32_t s, raw_len, nq = 0;
        if (p < e && text[p] == '"') {
            s = ++p;
            while (true) {
                if (p >= e) { csv_fail(err, r, DFM_CSV_ERR_QUOTE); return; }
                if (text[p] == '"') {
                    if (p + 1 < e && text[p + 1] == '"') { ++nq; p += 2; continue; }
                    break;
                }
                ++p;
            }
            raw_len = p - s;
            ++p;                                   // closing quote
            if (p < e && text[p] != ',') { csv_fail(err, r, DFM_CSV_ERR_QUOTE); return; }
        } else {
            s = p;
            while (p < e && text[p] != ',') {
                if (text[p] == '"') { csv_fail(err, r, DFM_CSV_ERR_QUOTE); return; }
                ++p;
            }
            raw_len = p - s;
        }
        if (field >= cfg.n_fields) { csv_fail(err, r, DFM_CSV_ERR_FIELDS); return; }
        const int kind = cfg.kind[field];
        if (kind == DFM_CSV_INT32) {
            int32_t v = cfg.int_default[field];
            if (raw_len) {
                uint32_t i = s, end = s + raw_len;
                while (i < end && text[i] == ' ') ++i;
                bool neg = false;
                if (i < end && (text[i] == '-' || text[i] == '+')) { neg = text[i] == '-'; ++i; }
                long long acc = 0;
                int nd = 0;
                while (i < end && text[i] >= '0' && text[i] <= '9') {
                    acc = acc * 10 + (text[i] - '0');
                    if (acc > 2147483648ll) acc = 2147483649ll;     // saturate: flagged below
                    ++i; ++nd;
                }
                while (i < end && text[i] == ' ') ++i;
                if (neg) acc = -acc;
                if (!nd || i != end || nq || acc > 2147483647ll || acc < -2147483648ll) { csv_fail(err, r, DFM_CSV_ERR_INT); return; }
                v = (int32_t)acc;
            }
            const int sl = cfg.slot[field];
            if (sl >= 0) int_out[sl][r] = v;
            if (field == cfg.label_field) labels[r] = v >= cfg.label_min ? 1.f : 0.f;
        } else if (kind == DFM_CSV_STRING) {
            const int sl = cfg.slot[field];
            if (raw_len) {
                str_pos[sl][r] = s;
                str_len[sl][r] = raw_len - nq;
                str_flag[sl][r] = nq ? 1 : 0;
            } else {
                str_pos[sl][r] = 0;
                str_len[sl][r] = (uint32_t)(cfg.def_off[field + 1] - cfg.def_off[field]);
                str_flag[sl][r] = 2;
            }
        }
        ++field;
        if (p >= e) break;
        ++p;                                       // the comma; a trailing comma leaves one more (empty) field
    }
    if (field != cfg.n_fields) csv_fail(err, r, DFM_CSV_ERR_FIELDS);
}

__global__ void __launch_bounds__(256) csv_copy_strings_kernel(const char* __restrict__ text, uint32_t n_rec, int n_str, CsvDev cfg,
                                                               const int32_t* __restrict__ str_field, const uint32_t* const* __restrict__ str_pos,
                                                               const uint32_t* const* __restrict__ offsets, const uint8_t* const* __restrict__ str_flag,
                                                               char* const* __restrict__ bytes_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rec * (uint32_t)n_str) return;
    const uint32_t sl = t / n_rec, r = t % n_rec;          // consecutive threads -> consecutive records of one column
    const uint32_t o0 = offsets[sl][r], len = offsets[sl][r + 1] - o0;
    char* dst = bytes_out[sl] + o0;
    const uint8_t fl = str_flag[sl][r];
    if (fl == 2) {
        const char* d = cfg.def_bytes + cfg.def_off[str_field[sl]];
        for (uint32_t i = 0; i < len; ++i) dst[i] = d[i];
    } else if (fl == 0) {
        const char* src = text + str_pos[sl][r];
        for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
    } else {
        const char* src = text + str_pos[sl][r];
        for (uint32_t i = 0; i < len; ++i) {
            const char c = *src++;
            if (c == '"') ++src;                               // '""' -> '"'
            dst[i] = c;
        }
    }
}

}  // namespace

struct dfm_csv_reader {
    int device = 0, n_fields = 0, n_int = 0, n_str = 0, label_field = -1, label_min = 0;
    int32_t max_records = 0;
    int64_t max_bytes = 0;
    std::vector<int32_t> kind, slot, str_field;
    cudaStream_t stream = nullptr;
    // device
    char* text = nullptr;                  // staging for dfm_csv_decode_host (max_bytes, 16-byte padded)
    int32_t *d_kind = nullptr, *d_slot = nullptr, *d_int_default = nullptr, *d_def_off = nullptr, *d_str_field = nullptr;
    char* d_def_bytes = nullptr;
    uint32_t *counts = nullptr, *line_start = nullptr;
    void* scan_temp = nullptr;
    uint32_t* d_total = nullptr;
    unsigned long long* d_err = nullptr;
    std::vector<int32_t*> int_out;
    std::vector<uint32_t*> str_pos, str_off;
    std::vector<uint8_t*> str_flag;
    std::vector<char*> str_bytes;
    int32_t** d_int_out = nullptr; uint32_t **d_str_pos = nullptr, **d_str_off = nullptr; uint8_t** d_str_flag = nullptr; char** d_str_bytes = nullptr;
    float* labels = nullptr;
    uint32_t* h_total = nullptr; unsigned long long* h_err = nullptr;      // pinned
    int32_t n_records = 0;
    // file-resident mode (dfm_csv_load): the whole text and its line starts stay in HBM, batches are lists of line numbers
    char* file_text = nullptr; uint32_t* file_line_start = nullptr; int64_t file_lines = 0, file_bytes = 0;
    int32_t* d_line_idx = nullptr; int32_t* h_line_idx = nullptr;     // [max_records], pinned staging
    std::string error;
};

static thread_local std::string g_csv_error;

#define CSV_FAIL(r, code, ...)                                   \
    do {                                                         \
        char buf_[256];                                          \
        snprintf(buf_, sizeof buf_, __VA_ARGS__);                \
        if (r) (r)->error = buf_; else g_csv_error = buf_;       \
        return code;                                             \
    } while (0)
#define CSV_CK(r, call)                                                                               \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) CSV_FAIL(r, DFM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));    \
    } while (0)

template <typename T>
static cudaError_t csv_alloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)); }

extern "C" const char* dfm_csv_last_error(const dfm_csv_reader* r) { return r ? r->error.c_str() : g_csv_error.c_str(); }

extern "C" void dfm_csv_destroy(dfm_csv_reader* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    void* ptrs[] = {r->text, r->d_kind, r->d_slot, r->d_int_default, r->d_def_off, r->d_str_field, r->d_def_bytes, r->counts, r->line_start,
                    r->scan_temp, r->d_total, r->d_err, r->d_int_out, r->d_str_pos, r->d_str_off, r->d_str_flag, r->d_str_bytes, r->labels};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto p : r->int_out) cudaFree(p);
    for (auto p : r->str_pos) cudaFree(p);
    for (auto p : r->str_off) cudaFree(p);
    for (auto p : r->str_flag) cudaFree(p);
    for (auto p : r->str_bytes) cudaFree(p);
    if (r->file_text) cudaFree(r->file_text);
    if (r->file_line_start) cudaFree(r->file_line_start);
    if (r->d_line_idx) cudaFree(r->d_line_idx);
    if (r->h_line_idx) cudaFreeHost(r->h_line_idx);
    if (r->h_total) cudaFreeHost(r->h_total);
    if (r->h_err) cudaFreeHost(r->h_err);
    delete r;
}

extern "C" int dfm_csv_create(const dfm_csv_config* cfg, dfm_csv_reader** out) {
    if (!cfg || !out || cfg->n_fields <= 0 || !cfg->kind || cfg->max_records <= 0 || cfg->max_bytes <= 0)
        CSV_FAIL((dfm_csv_reader*)nullptr, DFM_ERR_INVALID_ARG, "dfm_csv_create: bad configuration");
    if (cfg->max_bytes >= (int64_t)1 << 32) CSV_FAIL((dfm_csv_reader*)nullptr, DFM_ERR_UNSUPPORTED, "dfm_csv_create: max_bytes must be < 4 GiB");
    dfm_csv_reader* r = new dfm_csv_reader();
    r->device = cfg->device; r->n_fields = cfg->n_fields; r->max_records = cfg->max_records;
    r->max_bytes = (cfg->max_bytes + 15) / 16 * 16;
    r->label_field = cfg->label_field; r->label_min = cfg->label_min;
    std::vector<int32_t> int_default(cfg->n_fields, 0), def_off(cfg->n_fields + 1, 0);
    std::string def_bytes;
    for (int f = 0; f < cfg->n_fields; ++f) {
        const int k = cfg->kind[f];
        r->kind.push_back(k);
        if (k == DFM_CSV_INT32) {
            r->slot.push_back(r->n_int++);
            if (cfg->int_default) int_default[f] = cfg->int_default[f];
        } else if (k == DFM_CSV_STRING) {
            r->slot.push_back(r->n_str++);
            r->str_field.push_back(f);
            if (cfg->str_default && cfg->str_default[f]) def_bytes += cfg->str_default[f];
        } else if (k == DFM_CSV_SKIP) {
            r->slot.push_back(-1);
        } else {
            delete r;
            CSV_FAIL((dfm_csv_reader*)nullptr, DFM_ERR_INVALID_ARG, "dfm_csv_create: field %d has unknown kind %d", f, k);
        }
        def_off[f + 1] = (int32_t)def_bytes.size();
    }
    if (r->label_field >= 0 && (r->label_field >= cfg->n_fields || r->kind[r->label_field] != DFM_CSV_INT32)) {
        delete r;
        CSV_FAIL((dfm_csv_reader*)nullptr, DFM_ERR_INVALID_ARG, "dfm_csv_create: label_field must be an int32 field");
    }
    *out = r;     // from here on dfm_csv_destroy cleans up
#define CR(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) { g_csv_error = std::string(#call ": ") + cudaGetErrorString(e_); dfm_csv_destroy(r); *out = nullptr; return DFM_ERR_CUDA; } \
    } while (0)
    CR(cudaSetDevice(r->device));
    CR(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    const size_t nf = cfg->n_fields, mr = (size_t)r->max_records, n_chunks = (size_t)r->max_bytes / CSV_CHUNK;
    CR(csv_alloc(&r->text, (size_t)r->max_bytes));
    CR(csv_alloc(&r->d_kind, nf)); CR(csv_alloc(&r->d_slot, nf)); CR(csv_alloc(&r->d_int_default, nf)); CR(csv_alloc(&r->d_def_off, nf + 1));
    CR(csv_alloc(&r->d_str_field, (size_t)r->n_str)); CR(csv_alloc(&r->d_def_bytes, def_bytes.size()));
    CR(cudaMemcpy(r->d_kind, r->kind.data(), nf * 4, cudaMemcpyHostToDevice));
    CR(cudaMemcpy(r->d_slot, r->slot.data(), nf * 4, cudaMemcpyHostToDevice));
    CR(cudaMemcpy(r->d_int_default, int_default.data(), nf * 4, cudaMemcpyHostToDevice));
    CR(cudaMemcpy(r->d_def_off, def_off.data(), (nf + 1) * 4, cudaMemcpyHostToDevice));
    if (r->n_str) CR(cudaMemcpy(r->d_str_field, r->str_field.data(), (size_t)r->n_str * 4, cudaMemcpyHostToDevice));
    if (!def_bytes.empty()) CR(cudaMemcpy(r->d_def_bytes, def_bytes.data(), def_bytes.size(), cudaMemcpyHostToDevice));
    CR(csv_alloc(&r->counts, n_chunks + 1)); CR(csv_alloc(&r->line_start, mr + 2));
    CR(cudaMalloc(&r->scan_temp, std::max(prims::scan_temp_bytes((int64_t)n_chunks, 4), prims::scan_temp_bytes((int64_t)mr + 1, 4)) + 256));
    CR(csv_alloc(&r->d_total, 1)); CR(csv_alloc(&r->d_err, 1)); CR(csv_alloc(&r->labels, mr));
    CR(cudaMallocHost(&r->h_total, 4)); CR(cudaMallocHost(&r->h_err, 8));
    for (int i = 0; i < r->n_int; ++i) { int32_t* p = nullptr; CR(csv_alloc(&p, mr)); r->int_out.push_back(p); }
    for (int i = 0; i < r->n_str; ++i) {
        uint32_t *a = nullptr, *b = nullptr; uint8_t* c = nullptr; char* d = nullptr;
        CR(csv_alloc(&a, mr)); r->str_pos.push_back(a);
        CR(csv_alloc(&b, mr + 1)); r->str_off.push_back(b);
        CR(csv_alloc(&c, mr)); r->str_flag.push_back(c);
        // unquoting only shrinks a field; a default can be longer than the (empty) field it replaces
        const size_t def_len = (size_t)(def_off[r->str_field[i] + 1] - def_off[r->str_field[i]]);
        CR(csv_alloc(&d, (size_t)r->max_bytes + def_len * mr)); r->str_bytes.push_back(d);
    }
    CR(csv_alloc(&r->d_int_out, (size_t)r->n_int)); CR(csv_alloc(&r->d_str_pos, (size_t)r->n_str)); CR(csv_alloc(&r->d_str_off, (size_t)r->n_str));
    CR(csv_alloc(&r->d_str_flag, (size_t)r->n_str)); CR(csv_alloc(&r->d_str_bytes, (size_t)r->n_str));
    if (r->n_int) CR(cudaMemcpy(r->d_int_out, r->int_out.data(), (size_t)r->n_int * 8, cudaMemcpyHostToDevice));
    if (r->n_str) {
        CR(cudaMemcpy(r->d_str_pos, r->str_pos.data(), (size_t)r->n_str * 8, cudaMemcpyHostToDevice));
        CR(cudaMemcpy(r->d_str_off, r->str_off.data(), (size_t)r->n_str * 8, cudaMemcpyHostToDevice));
        CR(cudaMemcpy(r->d_str_flag, r->str_flag.data(), (size_t)r->n_str * 8, cudaMemcpyHostToDevice));
        CR(cudaMemcpy(r->d_str_bytes, r->str_bytes.data(), (size_t)r->n_str * 8, cudaMemcpyHostToDevice));
    }
#undef CR
    return DFM_OK;
}

// newline scan of `text_dev` -> line_start[0..n_rec] (device), n_rec on the host (one stream sync)
static int csv_split_lines(dfm_csv_reader* r, const char* text_dev, int64_t n_bytes, uint32_t* counts, void* scan_temp, uint32_t* line_start,
                           int64_t max_lines, int64_t* n_rec_out, cudaStream_t st) {
    const int64_t n_chunks = (n_bytes + CSV_CHUNK - 1) / CSV_CHUNK;
    const uint4* t16 = reinterpret_cast<const uint4*>(text_dev);
    csv_count_nl_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, st>>>(t16, n_bytes, counts);
    prims::exclusive_scan_u32(counts, counts, n_chunks, scan_temp, r->d_total, st, nullptr);
    csv_line_starts_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, st>>>(t16, n_bytes, counts, line_start, (uint32_t)std::min<int64_t>(max_lines + 1, 0xffffffffll));
    CSV_CK(r, cudaMemcpyAsync(r->h_total, r->d_total, 4, cudaMemcpyDeviceToHost, st));
    char last = 0;
    CSV_CK(r, cudaMemcpyAsync(&last, text_dev + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    CSV_CK(r, cudaStreamSynchronize(st));
    int64_t n_rec = *r->h_total;
    if (last != '\n') {                    // final record without a newline
        ++n_rec;
        if (n_rec <= max_lines) {
            const uint32_t end = (uint32_t)n_bytes;
            CSV_CK(r, cudaMemcpyAsync(line_start + n_rec, &end, 4, cudaMemcpyHostToDevice, st));
        }
    }
    *n_rec_out = n_rec;
    return DFM_OK;
}

// fields of n_rec records (record i = line line_idx[i], or line i) -> the reader's output columns; one stream sync
static int csv_parse_records(dfm_csv_reader* r, const char* text_dev, const uint32_t* line_start, const int32_t* line_idx, int64_t n_rec,
                             int32_t* n_records_out, cudaStream_t st) {
    CSV_CK(r, cudaMemsetAsync(r->d_err, 0xff, 8, st));
    CsvDev cfg{r->n_fields, r->d_kind, r->d_slot, r->d_int_default, r->d_def_off, r->d_def_bytes, r->label_field, r->label_min};
    csv_parse_kernel<<<(unsigned)((n_rec + 127) / 128), 128, 0, st>>>(text_dev, line_start, line_idx, (uint32_t)n_rec, cfg, r->d_int_out, r->d_str_pos,
                                                                      r->d_str_off, r->d_str_flag, r->labels, r->d_err);
    for (int i = 0; i < r->n_str; ++i)
        prims::exclusive_scan_u32(r->str_off[i], r->str_off[i], n_rec, r->scan_temp, r->str_off[i] + n_rec, st, nullptr);
    if (r->n_str) {
        const int64_t nt = n_rec * r->n_str;
        csv_copy_strings_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(text_dev, (uint32_t)n_rec, r->n_str, cfg, r->d_str_field, r->d_str_pos, r->d_str_off,
                                                                              r->d_str_flag, r->d_str_bytes);
    }
    CSV_CK(r, cudaMemcpyAsync(r->h_err, r->d_err, 8, cudaMemcpyDeviceToHost, st));
    CSV_CK(r, cudaStreamSynchronize(st));
    CSV_CK(r, cudaGetLastError());
    if (*r->h_err != ~0ull) {
        const unsigned code = (unsigned)(*r->h_err & 0xff);
        const unsigned long long rec = *r->h_err >> 8;
        const char* what = code == DFM_CSV_ERR_FIELDS ? "wrong number of fields" : code == DFM_CSV_ERR_INT ? "field is not a valid int32" : "quoting error";
        CSV_FAIL(r, DFM_ERR_PARSE, "record %llu: %s", rec, what);
    }
    r->n_records = (int32_t)n_rec;
    if (n_records_out) *n_records_out = (int32_t)n_rec;
    return DFM_OK;
}

static int csv_decode_impl(dfm_csv_reader* r, const char* text_dev, int64_t n_bytes, int32_t* n_records_out, cudaStream_t st) {
    if (n_bytes < 0 || n_bytes > r->max_bytes) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode: %lld bytes exceed max_bytes", (long long)n_bytes);
    if (reinterpret_cast<uintptr_t>(text_dev) & 15) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode: text must be 16-byte aligned (and readable up to the next multiple of 16)");
    r->n_records = 0;
    if (n_records_out) *n_records_out = 0;
    if (n_bytes == 0) return DFM_OK;
    int64_t n_rec = 0;
    int rc = csv_split_lines(r, text_dev, n_bytes, r->counts, r->scan_temp, r->line_start, r->max_records, &n_rec, st);
    if (rc) return rc;
    if (n_rec > r->max_records) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode: %lld records exceed max_records %d", (long long)n_rec, r->max_records);
    return csv_parse_records(r, text_dev, r->line_start, nullptr, n_rec, n_records_out, st);
}

// ---- file-resident mode ------------------------------------------------------------------------------------------
// A training CSV is small next to HBM (ML-100K train.csv: 13 MB), so the whole text is uploaded once and split into
// lines on the device; a batch is then just the list of line numbers the host's shuffle selected (256 KB for 65 536
// records instead of 10 MB of text), and no byte of the file is touched by the CPU again.
extern "C" int dfm_csv_load(dfm_csv_reader* r, const char* text_host, int64_t n_bytes, int64_t* n_lines_out) {
    if (!r || !text_host || n_bytes <= 0) return DFM_ERR_INVALID_ARG;
    if (n_bytes >= (int64_t)1 << 32) CSV_FAIL(r, DFM_ERR_UNSUPPORTED, "dfm_csv_load: files of 4 GiB and more are not supported");
    CSV_CK(r, cudaSetDevice(r->device));
    cudaStream_t st = r->stream;
    if (r->file_text) { cudaFree(r->file_text); r->file_text = nullptr; }
    if (r->file_line_start) { cudaFree(r->file_line_start); r->file_line_start = nullptr; }
    const int64_t padded = (n_bytes + 15) / 16 * 16, n_chunks = padded / CSV_CHUNK;
    CSV_CK(r, csv_alloc(&r->file_text, (size_t)padded));
    CSV_CK(r, cudaMemsetAsync(r->file_text + padded - 16, 0, 16, st));
    CSV_CK(r, cudaMemcpyAsync(r->file_text, text_host, (size_t)n_bytes, cudaMemcpyHostToDevice, st));
    uint32_t* counts = nullptr; void* scan_temp = nullptr; uint32_t* starts = nullptr;
    CSV_CK(r, csv_alloc(&counts, (size_t)n_chunks + 1));
    CSV_CK(r, cudaMalloc(&scan_temp, prims::scan_temp_bytes(n_chunks, 4) + 256));
    // first pass only counts the lines, so that the line-start array has the right size
    const uint4* t16 = reinterpret_cast<const uint4*>(r->file_text);
    csv_count_nl_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, st>>>(t16, n_bytes, counts);
    prims::exclusive_scan_u32(counts, counts, n_chunks, scan_temp, r->d_total, st, nullptr);
    CSV_CK(r, cudaMemcpyAsync(r->h_total, r->d_total, 4, cudaMemcpyDeviceToHost, st));
    CSV_CK(r, cudaStreamSynchronize(st));
    const int64_t cap = (int64_t)*r->h_total + 1;
    CSV_CK(r, csv_alloc(&starts, (size_t)cap + 2));
    int64_t n_lines = 0;
    int rc = csv_split_lines(r, r->file_text, n_bytes, counts, scan_temp, starts, cap, &n_lines, st);
    CSV_CK(r, cudaStreamSynchronize(st));
    cudaFree(counts); cudaFree(scan_temp);
    if (rc) { cudaFree(starts); return rc; }
    r->file_line_start = starts; r->file_lines = n_lines; r->file_bytes = n_bytes;
    if (!r->d_line_idx) {
        CSV_CK(r, csv_alloc(&r->d_line_idx, (size_t)r->max_records));
        CSV_CK(r, cudaMallocHost(&r->h_line_idx, (size_t)r->max_records * 4));
    }
    if (n_lines_out) *n_lines_out = n_lines;
    return DFM_OK;
}

extern "C" int dfm_csv_decode_lines(dfm_csv_reader* r, const int32_t* line_idx_host, int32_t n, void* stream) {
    if (!r || (!line_idx_host && n)) return DFM_ERR_INVALID_ARG;
    if (!r->file_text) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode_lines: call dfm_csv_load first");
    if (n < 0 || n > r->max_records) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode_lines: %d records exceed max_records %d", n, r->max_records);
    CSV_CK(r, cudaSetDevice(r->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : r->stream;
    r->n_records = 0;
    if (n == 0) return DFM_OK;
    for (int32_t i = 0; i < n; ++i) {
        if (line_idx_host[i] < 0 || line_idx_host[i] >= r->file_lines) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode_lines: line %d out of range", line_idx_host[i]);
        r->h_line_idx[i] = line_idx_host[i];
    }
    CSV_CK(r, cudaMemcpyAsync(r->d_line_idx, r->h_line_idx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    return csv_parse_records(r, r->file_text, r->file_line_start, r->d_line_idx, n, nullptr, st);
}

extern "C" int dfm_csv_decode(dfm_csv_reader* r, const char* text_dev, int64_t n_bytes, int32_t* n_records_out, void* stream) {
    if (!r || (!text_dev && n_bytes)) return DFM_ERR_INVALID_ARG;
    CSV_CK(r, cudaSetDevice(r->device));
    return csv_decode_impl(r, text_dev, n_bytes, n_records_out, stream ? (cudaStream_t)stream : r->stream);
}

extern "C" int dfm_csv_decode_host(dfm_csv_reader* r, const char* text_host, int64_t n_bytes, int32_t* n_records_out, void* stream) {
    if (!r || (!text_host && n_bytes)) return DFM_ERR_INVALID_ARG;
    if (n_bytes < 0 || n_bytes > r->max_bytes) CSV_FAIL(r, DFM_ERR_INVALID_ARG, "dfm_csv_decode_host: %lld bytes exceed max_bytes", (long long)n_bytes);
    CSV_CK(r, cudaSetDevice(r->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : r->stream;
    if (n_bytes) CSV_CK(r, cudaMemcpyAsync(r->text, text_host, (size_t)n_bytes, cudaMemcpyHostToDevice, st));
    return csv_decode_impl(r, r->text, n_bytes, n_records_out, st);
}

extern "C" int32_t dfm_csv_num_records(const dfm_csv_reader* r) { return r ? r->n_records : -1; }
extern "C" const int32_t* dfm_csv_int_column(const dfm_csv_reader* r, int32_t field) {
    if (!r || field < 0 || field >= r->n_fields || r->kind[field] != DFM_CSV_INT32) return nullptr;
    return r->int_out[r->slot[field]];
}
extern "C" const char* dfm_csv_str_bytes(const dfm_csv_reader* r, int32_t field) {
    if (!r || field < 0 || field >= r->n_fields || r->kind[field] != DFM_CSV_STRING) return nullptr;
    return r->str_bytes[r->slot[field]];
}
extern "C" const int32_t* dfm_csv_str_offsets(const dfm_csv_reader* r, int32_t field) {
    if (!r || field < 0 || field >= r->n_fields || r->kind[field] != DFM_CSV_STRING) return nullptr;
    return reinterpret_cast<const int32_t*>(r->str_off[r->slot[field]]);
}
extern "C" const float* dfm_csv_labels(const dfm_csv_reader* r) { return r ? r->labels : nullptr; }

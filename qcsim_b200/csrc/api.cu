// api.cu -- C ABI of libqcsim_b200.so (see include/qcsim_b200.h for the contract and the
// reference members each entry point replaces).  Host logic only: argument checks with the
// reference's error conventions, gate classification, kernel launches on the handle's stream.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/qcsim_b200.h"
#include "engine.h"
#include "dist.h"

using namespace qcsim;

namespace qcsim {
thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace qcsim

#define API_GUARD(h)                                                            \
  if (!(h)) return fail(QCSIM_ERR_BAD_ARG, "null register handle");             \
  {                                                                             \
    cudaError_t e__ = cudaSetDevice((h)->device);                               \
    if (e__ != cudaSuccess) return fail(QCSIM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e__)); \
  }

// A handle made by qcsim_sv_create_multi is a front for one shard per device (multi.cu): the call is repeated on every
// shard, each from its own worker thread; `s` is the shard, `r` its rank.  Value outputs are taken from rank 0 (every
// rank computes the same collectively reduced value).
#define FRONT(h, expr)                                                                  \
  if ((h)->multi) return multi_forward((h), [&](qcsim_sv* s, int r) -> int {            \
      (void)r;                                                                          \
      return (expr);                                                                    \
    })

extern "C" {

const char* qcsim_last_error(void) { return g_last_error.c_str(); }
int qcsim_abi_version(void) { return QCSIM_ABI_VERSION; }

int qcsim_device_count(int* count) {
  if (!count) return fail(QCSIM_ERR_BAD_ARG, "null count");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(QCSIM_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return QCSIM_OK;
}

int qcsim_sv_create(qcsim_sv** out, int n_qubits, int device) {
  return engine_create(out, n_qubits, device, 0, 1, nullptr);
}

int qcsim_nccl_unique_id(void* out_128_bytes) { return engine_nccl_unique_id(out_128_bytes); }

int qcsim_sv_create_sharded(qcsim_sv** out, int n_qubits, int device, int rank, int world, const void* nccl_id) {
  return engine_create(out, n_qubits, device, rank, world, nccl_id);
}

int qcsim_sv_create_multi(qcsim_sv** out, int n_qubits, int n_devices, const int* device_ids) {
  return multi_create(out, n_qubits, n_devices, device_ids);
}

int qcsim_sv_destroy(qcsim_sv* h) {
  if (!h) return QCSIM_OK;
  if (h->multi) return multi_destroy(h);
  cudaSetDevice(h->device);
  return engine_destroy(h);
}

int qcsim_sv_clone(const qcsim_sv* src, qcsim_sv** out) {
  if (!src || !out) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  if (src->multi) return fail(QCSIM_ERR_UNSUPPORTED, "clone of a multi-device register is not supported");
  return engine_clone(src, out);
}

int qcsim_sv_sync(qcsim_sv* h) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_sync(s));
  QCSIM_TRY(engine_flush(h));
  QCSIM_TRY(engine_wait(h));
  return QCSIM_OK;
}

int qcsim_sv_flush(qcsim_sv* h) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_flush(s));
  return engine_flush(h);
}

int qcsim_sv_n_qubits(const qcsim_sv* h, int* n_qubits, int* n_local_qubits) {
  if (!h) return fail(QCSIM_ERR_BAD_ARG, "null register handle");
  if (n_qubits) *n_qubits = h->n;
  if (n_local_qubits) *n_local_qubits = h->n_local;
  return QCSIM_OK;
}

int qcsim_sv_device_ptr(qcsim_sv* h, void** dptr, void** cuda_stream) {
  API_GUARD(h);
  if (h->multi) return fail(QCSIM_ERR_UNSUPPORTED, "a multi-device register has no single device pointer");
  QCSIM_TRY(engine_flush(h));
  QCSIM_TRY(engine_canonicalize(h));
  if (dptr) *dptr = h->psi;
  if (cuda_stream) *cuda_stream = (void*)h->stream;
  return QCSIM_OK;
}

/* ---- state setters / getters ---------------------------------------------------------------- */

int qcsim_sv_set_basis_state(qcsim_sv* h, uint64_t state) {
  API_GUARD(h);
  if (state >= h->dim) return fail(QCSIM_ERR_BAD_STATE, "basis state out of range");
  FRONT(h, qcsim_sv_set_basis_state(s, state));
  engine_drop_queue(h);
  return engine_set_basis_state(h, state);
}

int qcsim_sv_fill(qcsim_sv* h, double re, double im) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_fill(s, re, im));
  engine_drop_queue(h);
  return engine_fill(h, re, im);
}

int qcsim_sv_set_amplitude(qcsim_sv* h, uint64_t state, double re, double im) {
  API_GUARD(h);
  if (state >= h->dim) return fail(QCSIM_ERR_BAD_STATE, "basis state out of range");
  FRONT(h, qcsim_sv_set_amplitude(s, state, re, im));
  QCSIM_TRY(engine_flush(h));
  return engine_set_amplitude(h, state, re, im);
}

int qcsim_sv_get_amplitude(qcsim_sv* h, uint64_t state, double* re_im) {
  API_GUARD(h);
  if (!re_im) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (state >= h->dim) {
    re_im[0] = re_im[1] = 0;
    return fail(QCSIM_ERR_BAD_STATE, "basis state out of range");
  }
  if (h->multi) {
    double scratch[kMaxWorld][2];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_get_amplitude(s, state, r == 0 ? re_im : scratch[r]); });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_get_amplitude(h, state, re_im);
}

int qcsim_sv_upload(qcsim_sv* h, const double* host, uint64_t first, uint64_t count) {
  API_GUARD(h);
  if (!host && count) return fail(QCSIM_ERR_BAD_ARG, "null host buffer");
  if (h->multi) {
    if (first + count > h->dim || first + count < first) return fail(QCSIM_ERR_BAD_ARG, "range is outside the register");
    return multi_forward(h, [&](qcsim_sv* s, int r) {  // every shard takes the part of the range that falls into its slice
      const uint64_t lo = std::max<uint64_t>(first, (uint64_t)r << s->n_local), hi = std::min<uint64_t>(first + count, (uint64_t)(r + 1) << s->n_local);
      const uint64_t base = (uint64_t)r << s->n_local;
      return hi > lo ? qcsim_sv_upload(s, host + 2 * (lo - first), lo, hi - lo) : qcsim_sv_upload(s, host, base, 0);
    });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_transfer(h, const_cast<double*>(host), first, count, /*to_device=*/true);
}

int qcsim_sv_download(qcsim_sv* h, double* host, uint64_t first, uint64_t count) {
  API_GUARD(h);
  if (!host && count) return fail(QCSIM_ERR_BAD_ARG, "null host buffer");
  if (h->multi) {
    if (first + count > h->dim || first + count < first) return fail(QCSIM_ERR_BAD_ARG, "range is outside the register");
    return multi_forward(h, [&](qcsim_sv* s, int r) {
      const uint64_t lo = std::max<uint64_t>(first, (uint64_t)r << s->n_local), hi = std::min<uint64_t>(first + count, (uint64_t)(r + 1) << s->n_local);
      const uint64_t base = (uint64_t)r << s->n_local;
      return hi > lo ? qcsim_sv_download(s, host + 2 * (lo - first), lo, hi - lo) : qcsim_sv_download(s, host, base, 0);
    });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_transfer(h, host, first, count, /*to_device=*/false);
}

int qcsim_sv_norm2(qcsim_sv* h, double* out) {
  API_GUARD(h);
  if (!out) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (h->multi) {
    double scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_norm2(s, r == 0 ? out : &scratch[r]); });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_masked_norm2(h, 0, 0, out);
}

int qcsim_sv_scale(qcsim_sv* h, double factor) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_scale(s, factor));
  QCSIM_TRY(engine_flush(h));
  return engine_scale(h, factor);
}

int qcsim_sv_normalize(qcsim_sv* h) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_normalize(s));
  QCSIM_TRY(engine_flush(h));
  double n2 = 0;
  QCSIM_TRY(engine_masked_norm2(h, 0, 0, &n2));
  const double norm = std::sqrt(n2);
  if (norm < 1E-20) return QCSIM_OK;  // QubitRegister.h:127
  return engine_scale(h, 1. / norm);
}

int qcsim_sv_save_state(qcsim_sv* h) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_save_state(s));
  QCSIM_TRY(engine_flush(h));
  return engine_save(h);
}

int qcsim_sv_restore_state(qcsim_sv* h, int destructive) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_restore_state(s, destructive));
  if (!h->saved) return QCSIM_OK;  // nothing saved: the reference leaves the register alone (QubitRegister.h:607,613); queued gates stay queued
  engine_drop_queue(h);            // the restored state overwrites whatever the queued gates would have produced
  return engine_restore(h, destructive != 0);
}

int qcsim_sv_inner_product(qcsim_sv* a, qcsim_sv* b, double* re_im) {
  API_GUARD(a);
  if (!b || !re_im) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  if (a->multi || b->multi) {
    if (!a->multi || !b->multi || multi_world(a) != multi_world(b)) return fail(QCSIM_ERR_BAD_ARG, "registers must have the same shape and devices");
    double scratch[kMaxWorld][2];
    return multi_forward(a, [&](qcsim_sv* s, int r) { return qcsim_sv_inner_product(s, multi_shard(b, r), r == 0 ? re_im : scratch[r]); });
  }
  QCSIM_TRY(engine_flush(a));
  QCSIM_TRY(engine_flush(b));
  return engine_inner_product(a, b, re_im);
}

/* ---- gates ------------------------------------------------------------------------------------ */

static int check_qubits(const qcsim_sv* h, int nq, uint64_t q, uint64_t c1, uint64_t c2) {
  // QubitRegister::CheckQubits, QubitRegister.h:677-690 (same order of tests, same messages)
  const uint64_t n = (uint64_t)h->n;
  if (nq < 1 || nq > 3) return fail(QCSIM_ERR_BAD_ARG, "gate must act on 1, 2 or 3 qubits");
  if (n <= q) return fail(QCSIM_ERR_QUBIT_TOO_HIGH, "Qubit number is too high");
  if (nq == 2) {
    if (n <= c1) return fail(QCSIM_ERR_CTRL_TOO_HIGH, "Controlling qubit number is too high");
    if (q == c1) return fail(QCSIM_ERR_SAME_QUBITS, "Qubit and controlling qubit are the same");
  } else if (nq == 3) {
    if (n <= c1 || n <= c2) return fail(QCSIM_ERR_CTRL_TOO_HIGH, "Controlling qubit number is too high");
    if (q == c1 || q == c2 || c1 == c2) return fail(QCSIM_ERR_SAME_QUBITS, "Qubits must be different");
  }
  return QCSIM_OK;
}

int qcsim_sv_apply(qcsim_sv* h, int nq, const double* m, int flags, uint64_t q, uint64_t c1, uint64_t c2) {
  API_GUARD(h);
  if (!m) return fail(QCSIM_ERR_BAD_ARG, "null matrix");
  QCSIM_TRY(check_qubits(h, nq, q, c1, c2));
  FRONT(h, qcsim_sv_apply(s, nq, m, flags, q, c1, c2));
  h->stats.gates_applied++;
  const Op op = classify(nq, m, flags, q, c1, c2);
  if (h->fusion) return engine_enqueue(h, op);
  return engine_apply_now(h, op);
}

int qcsim_sv_apply_batch(qcsim_sv* h, const qcsim_gate* gates, uint64_t count) {
  API_GUARD(h);
  if (!gates && count) return fail(QCSIM_ERR_BAD_ARG, "null gate list");
  for (uint64_t i = 0; i < count; ++i) QCSIM_TRY(check_qubits(h, gates[i].nq, gates[i].q, gates[i].c1, gates[i].c2));
  FRONT(h, qcsim_sv_apply_batch(s, gates, count));
  for (uint64_t i = 0; i < count; ++i) {
    const qcsim_gate& g = gates[i];
    h->stats.gates_applied++;
    QCSIM_TRY(engine_enqueue(h, classify(g.nq, g.m, g.flags, g.q, g.c1, g.c2)));
  }
  if (!h->fusion) return engine_flush(h);
  return QCSIM_OK;
}

int qcsim_sv_apply_operator(qcsim_sv* h, const double* m) {
  API_GUARD(h);
  if (!m) return fail(QCSIM_ERR_BAD_ARG, "null matrix");
  if (h->world > 1 || h->multi) return fail(QCSIM_ERR_UNSUPPORTED, "ApplyOperatorMatrix is not supported on a sharded register");
  if (h->n > QCSIM_MAX_OPERATOR_QUBITS)
    return fail(QCSIM_ERR_UNSUPPORTED, "ApplyOperatorMatrix: a dense operator on %d qubits needs %.0f GiB; limit is %d qubits", h->n,
                std::ldexp(16.0, 2 * h->n - 30), QCSIM_MAX_OPERATOR_QUBITS);
  QCSIM_TRY(engine_flush(h));
  h->stats.gates_applied++;
  return engine_apply_operator(h, m);
}

/* ---- circuit files ---------------------------------------------------------------------------- */

static const char kCircuitMagic[8] = {'Q', 'C', 'S', 'I', 'M', 'C', '1', '\0'};

int qcsim_circuit_save(const char* path, uint32_t n_qubits, const qcsim_circuit_gate* gates, uint64_t count) {
  if (!path || (!gates && count)) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(QCSIM_ERR_BAD_ARG, "cannot open %s for writing", path);
  const uint32_t reserved = 0;
  bool ok = std::fwrite(kCircuitMagic, 1, 8, f) == 8 && std::fwrite(&n_qubits, 4, 1, f) == 1 && std::fwrite(&reserved, 4, 1, f) == 1 &&
            std::fwrite(&count, 8, 1, f) == 1;
  for (uint64_t i = 0; ok && i < count; ++i) {
    const qcsim_circuit_gate& g = gates[i];
    if (g.nq < 1 || g.nq > 3) {
      std::fclose(f);
      return fail(QCSIM_ERR_BAD_ARG, "gate %llu acts on %d qubits", (unsigned long long)i, g.nq);
    }
    const size_t nm = (size_t)2 << (2 * g.nq);
    ok = std::fwrite(&g.nq, 4, 4, f) == 4 && std::fwrite(&g.q, 8, 3, f) == 3 && std::fwrite(g.params, 8, 4, f) == 4 && std::fwrite(g.m, 8, nm, f) == nm;
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? QCSIM_OK : fail(QCSIM_ERR_BAD_ARG, "short write to %s", path);
}

int qcsim_circuit_load(const char* path, uint32_t* n_qubits, qcsim_circuit_gate** gates, uint64_t* count) {
  if (!path || !gates || !count) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  *gates = nullptr;
  *count = 0;
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(QCSIM_ERR_BAD_ARG, "cannot open %s", path);
  char magic[8];
  uint32_t nq = 0, reserved = 0;
  uint64_t cnt = 0;
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, kCircuitMagic, 8) != 0 || std::fread(&nq, 4, 1, f) != 1 || std::fread(&reserved, 4, 1, f) != 1 ||
      std::fread(&cnt, 8, 1, f) != 1 || cnt > (1ULL << 32)) {
    std::fclose(f);
    return fail(QCSIM_ERR_BAD_ARG, "%s is not a qcsim circuit file", path);
  }
  qcsim_circuit_gate* out = static_cast<qcsim_circuit_gate*>(std::calloc(cnt ? cnt : 1, sizeof(qcsim_circuit_gate)));
  if (!out) {
    std::fclose(f);
    return fail(QCSIM_ERR_OOM, "out of host memory for %llu gates", (unsigned long long)cnt);
  }
  for (uint64_t i = 0; i < cnt; ++i) {
    qcsim_circuit_gate& g = out[i];
    bool ok = std::fread(&g.nq, 4, 4, f) == 4 && g.nq >= 1 && g.nq <= 3 && std::fread(&g.q, 8, 3, f) == 3 && std::fread(g.params, 8, 4, f) == 4;
    const size_t nm = ok ? (size_t)2 << (2 * g.nq) : 0;
    ok = ok && std::fread(g.m, 8, nm, f) == nm;
    if (!ok) {
      std::free(out);
      std::fclose(f);
      return fail(QCSIM_ERR_BAD_ARG, "%s: truncated or corrupt at gate %llu", path, (unsigned long long)i);
    }
  }
  std::fclose(f);
  if (n_qubits) *n_qubits = nq;
  *gates = out;
  *count = cnt;
  return QCSIM_OK;
}

void qcsim_circuit_free(qcsim_circuit_gate* gates) { std::free(gates); }

int qcsim_sv_apply_circuit_file(qcsim_sv* h, const char* path) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_apply_circuit_file(s, path));
  uint32_t nq = 0;
  qcsim_circuit_gate* gates = nullptr;
  uint64_t count = 0;
  QCSIM_TRY(qcsim_circuit_load(path, &nq, &gates, &count));
  int rc = QCSIM_OK;
  if ((int)nq != h->n) rc = fail(QCSIM_ERR_BAD_ARG, "%s was recorded for %u qubits, the register has %d", path, nq, h->n);
  for (uint64_t i = 0; rc == QCSIM_OK && i < count; ++i) rc = check_qubits(h, gates[i].nq, gates[i].q, gates[i].c1, gates[i].c2);
  for (uint64_t i = 0; rc == QCSIM_OK && i < count; ++i) {
    const qcsim_circuit_gate& g = gates[i];
    h->stats.gates_applied++;
    rc = engine_enqueue(h, classify(g.nq, g.m, g.flags, g.q, g.c1, g.c2));
  }
  qcsim_circuit_free(gates);
  if (rc == QCSIM_OK && !h->fusion) rc = engine_flush(h);
  return rc;
}

int qcsim_sv_set_fusion(qcsim_sv* h, int enabled) {
  API_GUARD(h);
  if (h->multi) h->fusion = enabled != 0;
  FRONT(h, qcsim_sv_set_fusion(s, enabled));
  if (!enabled) QCSIM_TRY(engine_flush(h));
  h->fusion = enabled != 0;
  return QCSIM_OK;
}

int qcsim_sv_qft(qcsim_sv* h, uint64_t sq, uint64_t eq, int do_swap, int inverse) {
  API_GUARD(h);
  FRONT(h, qcsim_sv_qft(s, sq, eq, do_swap, inverse));
  return engine_qft(h, sq, eq, do_swap != 0, inverse != 0);
}

/* ---- measurement ---------------------------------------------------------------------------- */

int qcsim_sv_measure_all(qcsim_sv* h, double prob, uint64_t* outcome) {
  API_GUARD(h);
  if (!outcome) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (h->multi) {
    uint64_t scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_measure_all(s, prob, r == 0 ? outcome : &scratch[r]); });
  }
  QCSIM_TRY(engine_flush(h));
  uint64_t s = 0;
  QCSIM_TRY(engine_pick_state(h, prob, h->dim - 1, &s));  // fallback: last state, QubitRegister.h:173
  QCSIM_TRY(engine_set_basis_state(h, s));                // collapse, :192
  *outcome = s;
  return QCSIM_OK;
}

int qcsim_sv_measure_all_nocollapse(qcsim_sv* h, double prob, uint64_t* outcome) {
  API_GUARD(h);
  if (!outcome) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (h->multi) {
    uint64_t scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_measure_all_nocollapse(s, prob, r == 0 ? outcome : &scratch[r]); });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_pick_state(h, prob, 0, outcome);  // fallback 0, QubitRegister.h:623
}

static int measured_mask(const qcsim_sv* h, uint64_t first, uint64_t last, uint64_t* mask) {
  if (first > last || last >= (uint64_t)h->n) return fail(QCSIM_ERR_BAD_ARG, "bad measured qubit range");
  const uint64_t low = (1ULL << first) - 1ULL;
  const uint64_t upto = (last + 1 >= 64) ? ~0ULL : ((1ULL << (last + 1)) - 1ULL);
  *mask = upto - low;  // QubitRegisterCalculator.h:1128-1130
  return QCSIM_OK;
}

int qcsim_sv_measure(qcsim_sv* h, uint64_t first, uint64_t last, double prob, uint64_t* outcome) {
  API_GUARD(h);
  if (!outcome) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (h->multi) {
    uint64_t scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_measure(s, first, last, prob, r == 0 ? outcome : &scratch[r]); });
  }
  uint64_t mask = 0;
  QCSIM_TRY(measured_mask(h, first, last, &mask));
  QCSIM_TRY(engine_flush(h));
  uint64_t s = 0;
  QCSIM_TRY(engine_pick_state(h, prob, 0, &s));  // fallback 0, QubitRegisterCalculator.h:954,1132
  const uint64_t want = s & mask;
  double acc = 0;
  QCSIM_TRY(engine_masked_norm2(h, mask, want, &acc));
  QCSIM_TRY(engine_collapse(h, mask, want, 1. / std::sqrt(acc)));
  *outcome = want >> first;
  return QCSIM_OK;
}

int qcsim_sv_measure_nocollapse(qcsim_sv* h, uint64_t first, uint64_t last, double prob, uint64_t* outcome) {
  API_GUARD(h);
  if (!outcome) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (h->multi) {
    uint64_t scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_measure_nocollapse(s, first, last, prob, r == 0 ? outcome : &scratch[r]); });
  }
  uint64_t mask = 0;
  QCSIM_TRY(measured_mask(h, first, last, &mask));
  QCSIM_TRY(engine_flush(h));
  uint64_t s = 0;
  QCSIM_TRY(engine_pick_state(h, prob, 0, &s));
  *outcome = (s & mask) >> first;
  return QCSIM_OK;
}

int qcsim_sv_qubit_probability(qcsim_sv* h, uint64_t q, double* p) {
  API_GUARD(h);
  if (!p) return fail(QCSIM_ERR_BAD_ARG, "null output");
  if (q >= (uint64_t)h->n) return fail(QCSIM_ERR_QUBIT_TOO_HIGH, "Qubit number is too high");
  if (h->multi) {
    double scratch[kMaxWorld];
    return multi_forward(h, [&](qcsim_sv* s, int r) { return qcsim_sv_qubit_probability(s, q, r == 0 ? p : &scratch[r]); });
  }
  QCSIM_TRY(engine_flush_for_diagonal_observable(h, 1ULL << q));  // gates that cannot change P(q) stay queued
  return engine_masked_norm2(h, 1ULL << q, 1ULL << q, p);
}

int qcsim_sv_sample(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes) {
  API_GUARD(h);
  if ((!probs || !outcomes) && count) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  if (h->multi) {
    std::vector<std::vector<uint64_t>> scratch(multi_world(h));
    return multi_forward(h, [&](qcsim_sv* s, int r) {
      if (r != 0) scratch[r].resize(count);
      return qcsim_sv_sample(s, probs, count, r == 0 ? outcomes : scratch[r].data());
    });
  }
  QCSIM_TRY(engine_flush(h));
  return engine_sample(h, probs, count, outcomes);
}

int qcsim_sv_set_strict_measure(qcsim_sv* h, int enabled) {
  if (!h) return fail(QCSIM_ERR_BAD_ARG, "null register handle");
  h->strict_measure = enabled != 0;  // kept for ABI compatibility: measurements are always exact
  return QCSIM_OK;
}

int qcsim_sv_get_stats(const qcsim_sv* h, qcsim_stats* out) {
  if (!h || !out) return fail(QCSIM_ERR_BAD_ARG, "null argument");
  if (h->multi) {  // rank 0's counters (the shards run in lock step); its exchange timings resolved on its own thread
    qcsim_stats scratch[kMaxWorld];
    return multi_forward(const_cast<qcsim_sv*>(h), [&](qcsim_sv* s, int r) { return qcsim_sv_get_stats(s, r == 0 ? out : &scratch[r]); });
  }
  if (h->world > 1) dist_collect_stats(const_cast<qcsim_sv*>(h));
  *out = h->stats;
  return QCSIM_OK;
}

int qcsim_sv_reset_stats(qcsim_sv* h) {
  if (!h) return fail(QCSIM_ERR_BAD_ARG, "null register handle");
  FRONT(h, qcsim_sv_reset_stats(s));
  if (h->world > 1) dist_collect_stats(h);  // drop the timings of exchanges that finished before the reset
  std::memset(&h->stats, 0, sizeof(h->stats));
  return QCSIM_OK;
}

}  // extern "C"

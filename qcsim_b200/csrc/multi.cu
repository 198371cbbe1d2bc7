// multi.cu -- one host thread drives several GPUs: qcsim_sv_create_multi ("qcsim_sv_create(&h, n_qubits,
// n_devices, device_ids)").
//
// The drop-in class QC::QubitRegister is used from ONE thread of ONE process (QubitRegister.h:17-57), while the
// sharded engine (dist.cu) is SPMD: one rank per GPU, every rank makes the same calls.  This front closes the gap
// without a second code path: it owns one sharded register per device, each driven by a persistent worker thread
// bound to that device; every API call on the front is handed to all workers at once and returns when all are
// done.  The shards talk to each other exactly as separate processes would -- NCCL for the few scalars, the
// in-place exchange kernel over peer memory for qubit exchanges -- except that peers of the same process are mapped
// with cudaDeviceEnablePeerAccess instead of CUDA IPC (dist.cu: setup_peers).
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "dist.h"
#include "engine.h"

namespace qcsim {

namespace {

struct Multi {
  int world = 0;
  std::vector<int> devices;
  std::vector<qcsim_sv*> shards;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  const std::function<int(qcsim_sv*, int)>* job = nullptr;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;
  std::vector<std::string> err;
};

Multi* mm(const qcsim_sv* h) { return static_cast<Multi*>(h->multi); }

void worker_main(Multi* m, int rank) {
  cudaSetDevice(m->devices[rank]);
  uint64_t seen = 0;
  for (;;) {
    const std::function<int(qcsim_sv*, int)>* job = nullptr;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_job.wait(lk, [&] { return m->stop || m->generation != seen; });
      if (m->stop) return;
      seen = m->generation;
      job = m->job;
    }
    g_last_error.clear();
    const int rc = (*job)(m->shards[rank], rank);
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->rc[rank] = rc;
      m->err[rank] = g_last_error;
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}

int run_all(Multi* m, const std::function<int(qcsim_sv*, int)>& fn) {
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->job = &fn;
    m->pending = m->world;
    ++m->generation;
  }
  m->cv_job.notify_all();
  std::unique_lock<std::mutex> lk(m->mu);
  m->cv_done.wait(lk, [&] { return m->pending == 0; });
  for (int r = 0; r < m->world; ++r)
    if (m->rc[r] != QCSIM_OK) {
      g_last_error = m->err[r];
      return m->rc[r];
    }
  return QCSIM_OK;
}

}  // namespace

int multi_world(const qcsim_sv* front) { return mm(front)->world; }
qcsim_sv* multi_shard(const qcsim_sv* front, int rank) { return mm(front)->shards[rank]; }

int multi_forward(qcsim_sv* front, const std::function<int(qcsim_sv*, int)>& fn) { return run_all(mm(front), fn); }

int multi_create(qcsim_sv** out, int n_qubits, int n_devices, const int* device_ids) {
  if (!out) return fail(QCSIM_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if (n_devices < 1 || (n_devices & (n_devices - 1)) || n_devices > kMaxWorld)
    return fail(QCSIM_ERR_BAD_ARG, "n_devices must be a power of two in 1..%d", kMaxWorld);
  if (n_devices == 1) return engine_create(out, n_qubits, device_ids ? device_ids[0] : 0, 0, 1, nullptr);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(QCSIM_ERR_CUDA, "no CUDA device; qcsim_b200 has no CPU fallback");
  Multi* m = new Multi();
  m->world = n_devices;
  for (int r = 0; r < n_devices; ++r) {
    const int dev = device_ids ? device_ids[r] : r;
    if (dev < 0 || dev >= ndev) {
      delete m;
      return fail(QCSIM_ERR_BAD_ARG, "device %d out of range (%d devices)", dev, ndev);
    }
    for (int q = 0; q < r; ++q)
      if (m->devices[q] == dev) {
        delete m;
        return fail(QCSIM_ERR_BAD_ARG, "device %d listed twice", dev);
      }
    m->devices.push_back(dev);
  }
  m->shards.assign(n_devices, nullptr);
  m->rc.assign(n_devices, QCSIM_OK);
  m->err.assign(n_devices, std::string());
  unsigned char id[128];
  const int rc_id = engine_nccl_unique_id(id);
  if (rc_id != QCSIM_OK) {
    delete m;
    return rc_id;
  }
  for (int r = 0; r < n_devices; ++r) m->workers.emplace_back(worker_main, m, r);
  // every shard is created by its own worker: ncclCommInitRank is a collective over the N threads
  const int rc = run_all(m, [&](qcsim_sv*, int r) { return engine_create(&m->shards[r], n_qubits, m->devices[r], r, n_devices, id); });
  qcsim_sv* front = new qcsim_sv();
  front->multi = m;
  front->n = n_qubits;
  int log2w = 0;
  while ((1 << log2w) < n_devices) ++log2w;
  front->n_local = n_qubits - log2w;
  front->dim = 1ULL << n_qubits;
  front->dim_local = 1ULL << front->n_local;
  front->device = m->devices[0];
  front->world = 1;  // the front itself is not a rank
  if (rc != QCSIM_OK) {
    const std::string keep = g_last_error;
    multi_destroy(front);
    g_last_error = keep;
    return rc;
  }
  *out = front;
  return QCSIM_OK;
}

int multi_destroy(qcsim_sv* front) {
  Multi* m = mm(front);
  run_all(m, [&](qcsim_sv* s, int r) {
    if (s) engine_destroy(s);
    m->shards[r] = nullptr;
    return QCSIM_OK;
  });
  {
    std::lock_guard<std::mutex> lk(m->mu);
    m->stop = true;
  }
  m->cv_job.notify_all();
  for (std::thread& t : m->workers) t.join();
  delete m;
  front->multi = nullptr;
  delete front;
  return QCSIM_OK;
}

}  // namespace qcsim

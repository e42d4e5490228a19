// Native input pipeline: gather (+ CIFAR augmentation) of uint8 image batches into caller-provided (pinned) staging memory,
// executed by a small pool of sleeping worker threads.
//
// Reference counterpart: the torch-0.3 DataLoader copy with `next_batch()` and its multiprocessing worker loop
// (src/data_loader_ops/my_data_loader.py:37-53,137-251) plus torchvision's RandomCrop(32, padding=4, reflect) +
// RandomHorizontalFlip transforms (src/util.py:37-52), all Python.  Here one job = one sub-batch: for every sample the
// source image is read once and written once, the reflect padding is never materialised (coordinates are mirrored on the
// fly) and the crop offsets / flip bits come from the caller, who draws them from a seeded generator so that every holder
// of a (step, batch) produces the same pixels -- the precondition of the exact-equality vote.
//
// Threads block on a condition variable when idle (no spinning: a GPU process that burns CPU runs into the container's
// CPU quota and gets descheduled, see utils/metrics.py).
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <set>
#include <thread>
#include <vector>

namespace {

struct Job {
  int ticket;
  std::vector<int64_t> idx;
  std::vector<int32_t> dx, dy;
  std::vector<uint8_t> flip;
  int pad;                 // < 0: plain gather
  uint8_t* out_images;
  int64_t* out_labels;
};

struct Loader {
  const uint8_t* images;
  const int64_t* labels;
  int64_t n_items;
  int C, H, W;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::deque<Job> queue;
  std::set<int> pending;   // tickets submitted and not yet finished
  int next_ticket = 0;
  int error = 0;
  bool stop = false;
};

inline int reflect(int q, int n) {
  if (q < 0) q = -q;
  if (q >= n) q = 2 * (n - 1) - q;
  return q;
}

int run_job(const Loader& L, const Job& j) {
  const int C = L.C, H = L.H, W = L.W;
  const size_t item = (size_t)C * H * W;
  const int n = (int)j.idx.size();
  for (int i = 0; i < n; ++i) {
    const int64_t src_i = j.idx[i];
    if (src_i < 0 || src_i >= L.n_items) return 1;
    const uint8_t* src = L.images + (size_t)src_i * item;
    uint8_t* dst = j.out_images + (size_t)i * item;
    if (j.out_labels && L.labels) j.out_labels[i] = L.labels[src_i];
    if (j.pad < 0) {
      std::memcpy(dst, src, item);
      continue;
    }
    // out[c, y, x] = padded[c, y + dy, (flip ? W-1-x : x) + dx], padded[p] = src[reflect(p - pad)]
    const int ox = j.dx[i] - j.pad, oy = j.dy[i] - j.pad;
    const bool fl = j.flip[i] != 0;
    int colmap[256];
    if (W > 256) return 2;
    for (int x = 0; x < W; ++x) colmap[x] = reflect((fl ? W - 1 - x : x) + ox, W);
    for (int c = 0; c < C; ++c) {
      for (int y = 0; y < H; ++y) {
        const uint8_t* srow = src + ((size_t)c * H + reflect(y + oy, H)) * W;
        uint8_t* drow = dst + ((size_t)c * H + y) * W;
        if (!fl && ox == 0) {
          std::memcpy(drow, srow, W);
        } else {
          for (int x = 0; x < W; ++x) drow[x] = srow[colmap[x]];
        }
      }
    }
  }
  return 0;
}

void worker_loop(Loader* L) {
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(L->mu);
      L->cv_job.wait(lk, [&] { return L->stop || !L->queue.empty(); });
      if (L->stop && L->queue.empty()) return;
      job = std::move(L->queue.front());
      L->queue.pop_front();
    }
    const int rc = run_job(*L, job);
    {
      std::lock_guard<std::mutex> lk(L->mu);
      if (rc) L->error = rc;
      L->pending.erase(job.ticket);
    }
    L->cv_done.notify_all();
  }
}

}  // namespace

extern "C" {

// Synchronous single-call forms (also the oracle for the threaded path).
int drc_host_gather_augment(const uint8_t* images, const int64_t* labels, int64_t n_items, int C, int H, int W, const int64_t* idx,
                            int n, const int32_t* dx, const int32_t* dy, const uint8_t* flip, int pad, uint8_t* out_images,
                            int64_t* out_labels) {
  Loader L;
  L.images = images; L.labels = labels; L.n_items = n_items; L.C = C; L.H = H; L.W = W;
  Job j;
  j.ticket = 0;
  j.idx.assign(idx, idx + n);
  if (pad >= 0) { j.dx.assign(dx, dx + n); j.dy.assign(dy, dy + n); j.flip.assign(flip, flip + n); }
  j.pad = pad; j.out_images = out_images; j.out_labels = out_labels;
  return run_job(L, j);
}

void* drc_loader_create(const uint8_t* images, const int64_t* labels, int64_t n_items, int C, int H, int W, int n_threads) {
  Loader* L = new Loader();
  L->images = images; L->labels = labels; L->n_items = n_items; L->C = C; L->H = H; L->W = W;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 16) n_threads = 16;
  for (int t = 0; t < n_threads; ++t) L->threads.emplace_back(worker_loop, L);
  return L;
}

// Enqueue one sub-batch; returns a ticket (>= 0).  The index / offset arrays are copied, the output memory must stay valid
// until the ticket has been waited for.  pad < 0: plain gather (dx / dy / flip may be null).
int drc_loader_submit(void* handle, const int64_t* idx, int n, const int32_t* dx, const int32_t* dy, const uint8_t* flip, int pad,
                      uint8_t* out_images, int64_t* out_labels) {
  Loader* L = static_cast<Loader*>(handle);
  Job j;
  j.idx.assign(idx, idx + n);
  if (pad >= 0) { j.dx.assign(dx, dx + n); j.dy.assign(dy, dy + n); j.flip.assign(flip, flip + n); }
  j.pad = pad; j.out_images = out_images; j.out_labels = out_labels;
  int ticket;
  {
    std::lock_guard<std::mutex> lk(L->mu);
    ticket = j.ticket = L->next_ticket++;
    L->pending.insert(ticket);
    L->queue.push_back(std::move(j));
  }
  L->cv_job.notify_one();
  return ticket;
}

// Block until every ticket <= `ticket` has finished.  Returns 0, or the first error code seen (1: index out of range,
// 2: image too wide).
int drc_loader_wait(void* handle, int ticket) {
  Loader* L = static_cast<Loader*>(handle);
  std::unique_lock<std::mutex> lk(L->mu);
  L->cv_done.wait(lk, [&] { return L->pending.empty() || *L->pending.begin() > ticket; });
  const int e = L->error;
  L->error = 0;
  return e;
}

void drc_loader_destroy(void* handle) {
  Loader* L = static_cast<Loader*>(handle);
  {
    std::lock_guard<std::mutex> lk(L->mu);
    L->stop = true;
  }
  L->cv_job.notify_all();
  for (auto& t : L->threads) t.join();
  delete L;
}

}  // extern "C"

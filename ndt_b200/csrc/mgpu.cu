/*
 * mgpu.cu -- all GPUs of one box behind the C ABI (ndt_b200.h: ndt_b200_mgpu_*).
 *
 * The reference spreads a frame over MPI ranks by rows (row j of rank r, thread t when
 * j = r*T + t mod size*T, ndt.c:812-820) and an animation by frames (ndt.c:1771-1787), then
 * collects whole images on rank 0 with a software tree of image_add (mpi_collect_image,
 * ndt.c:1277-1309).  Pixels and frames are independent (ndt.c:750-757), so here
 *
 *   - one host thread per context, NDT_B200_CTX_PER_DEVICE (2) contexts per GPU -- two frames /
 *     bands in flight per GPU hide the tail of the persistent trace kernels;
 *   - the work items (row bands of one frame, or whole frames of an animation) are pulled from ONE
 *     atomic counter: a dynamic queue instead of the reference's static cyclic rows, so the expensive
 *     bands (the hypercube's silhouette) do not serialise behind a static assignment;
 *   - every GPU copies the tiles it rendered straight into the rows of the caller's host frame over
 *     its own PCIe link -- the "gather to rank 0" without a hop through GPU 0.  There is no data-path
 *     collective: nothing is exchanged between GPUs.  (With one PROCESS per GPU -- bench.py under
 *     torchrun -- the same gather is NCCL send/recv into rank 0's HBM: ndt_b200/multi.py.)
 *
 * Scenes are replicated: every context uploads the flat scene it is about to render (a few MB).
 * For an animation whose scene_setup carries state from frame to frame (scenes/balls.c:27,181) the
 * host produces the flat scenes in order and hands them over as they are ready
 * (ndt_b200_mgpu_submit); the frames are rendered by whichever context is free.
 */
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include "ndt_internal.h"

#ifndef NDT_B200_CTX_PER_DEVICE
#define NDT_B200_CTX_PER_DEVICE 2
#endif

namespace {

struct Job {                      /* one frame handed to ndt_b200_mgpu_submit */
    ndt_flat_scene *fs;           /* private copy (freed by the worker) */
    uint8_t *u8;
    double *f64;
};

struct Worker {
    int device;
    ndt_b200_ctx *ctx;
    std::thread th;
};

}  // namespace

struct ndt_b200_mgpu {
    std::vector<Worker> w;
    int n_devices;
    /* streaming queue (ndt_b200_mgpu_submit / _wait) */
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> jobs;
    int in_flight;                /* queued + being rendered */
    int max_queue;
    bool stop;
    int err;                      /* first failure of a streamed frame */
    std::string err_msg;
    ndt_b200_stats acc;
};

static void stats_sum(ndt_b200_stats *a, const ndt_b200_stats *s)
{
    a->rays_primary += s->rays_primary; a->rays_bounce += s->rays_bounce; a->rays_shadow += s->rays_shadow;
    a->rays_unique += s->rays_unique; a->rays_ref += s->rays_ref; a->samples += s->samples; a->flops += s->flops;
    a->launches += s->launches;
    if (s->generations > a->generations) a->generations = s->generations;
    a->device_ms += s->device_ms;
}

static void stream_worker(ndt_b200_mgpu *m, Worker *me)
{
    cudaSetDevice(me->device);
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_job.wait(lk, [&] { return m->stop || !m->jobs.empty(); });
            if (m->jobs.empty()) return;        /* stop */
            j = m->jobs.front();
            m->jobs.pop_front();
        }
        int r = 0;
        ndt_b200_stats st;
        memset(&st, 0, sizeof st);
        if (!m->err) {
            r = ndt_b200_upload(me->ctx, j.fs);
            if (!r) r = ndt_b200_render_tile(me->ctx, 0, 0, j.fs->h.width, j.fs->h.height, j.f64, j.u8, NULL, NULL, NULL, &st);
        }
        ndt_b200_free_flat(j.fs);
        {
            std::lock_guard<std::mutex> lk(m->mu);
            if (r && !m->err) { m->err = r; m->err_msg = ndt_b200_last_error(); }
            if (!r) stats_sum(&m->acc, &st);
            --m->in_flight;
        }
        m->cv_done.notify_all();
    }
}

extern "C" int ndt_b200_mgpu_init(int n_devices, const int *devices, ndt_b200_mgpu **out)
{
    if (!out) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_mgpu_init: NULL out");
    *out = NULL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ndt_set_error(NDT_B200_E_CUDA, "no CUDA device (%s); libndt_b200 has no CPU path",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (n_devices <= 0) { n_devices = ndev; devices = NULL; }
    if (n_devices > ndev && !devices) return ndt_set_error(NDT_B200_E_ARG, "%d devices asked, %d present", n_devices, ndev);
    ndt_b200_mgpu *m = new (std::nothrow) ndt_b200_mgpu();
    if (!m) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    m->n_devices = n_devices;
    m->in_flight = 0; m->stop = false; m->err = 0;
    memset(&m->acc, 0, sizeof m->acc);
    m->w.resize((size_t)n_devices * NDT_B200_CTX_PER_DEVICE);
    m->max_queue = 2 * (int)m->w.size();
    for (size_t i = 0; i < m->w.size(); ++i) { m->w[i].ctx = NULL; m->w[i].device = -1; }
    for (int d = 0; d < n_devices; ++d) {
        const int dev = devices ? devices[d] : d;
        for (int k = 0; k < NDT_B200_CTX_PER_DEVICE; ++k) {
            Worker &w = m->w[(size_t)d * NDT_B200_CTX_PER_DEVICE + k];
            w.device = dev;
            int r = ndt_b200_init(dev, &w.ctx);
            if (r) { ndt_b200_mgpu_destroy(m); return r; }
        }
    }
    for (size_t i = 0; i < m->w.size(); ++i) m->w[i].th = std::thread(stream_worker, m, &m->w[i]);
    *out = m;
    return 0;
}

extern "C" void ndt_b200_mgpu_destroy(ndt_b200_mgpu *m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->stop = true;
    }
    m->cv_job.notify_all();
    for (size_t i = 0; i < m->w.size(); ++i) if (m->w[i].th.joinable()) m->w[i].th.join();
    for (size_t i = 0; i < m->w.size(); ++i) if (m->w[i].ctx) ndt_b200_destroy(m->w[i].ctx);
    delete m;
}

extern "C" int ndt_b200_mgpu_devices(const ndt_b200_mgpu *m) { return m ? m->n_devices : 0; }

/* one frame: row bands from a shared counter */
extern "C" int ndt_b200_mgpu_render_frame(ndt_b200_mgpu *m, const ndt_flat_scene *fs, int band_rows,
                                          double *rgba_f64, uint8_t *rgba_u8, uint8_t *hit, int32_t *obj_id,
                                          double *inv_depth, ndt_b200_stats *stats)
{
    if (!m || !fs) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_mgpu_render_frame: NULL argument");
    const int W = fs->h.width, H = fs->h.height;
    const int nctx = (int)m->w.size();
    if (band_rows <= 0) {
        /* about four bands per context, whole 8x4 ray blocks per band */
        band_rows = (H + 4 * nctx - 1) / (4 * nctx);
        band_rows = (band_rows + 3) & ~3;
        if (band_rows < 4) band_rows = 4;
    }
    const int nbands = (H + band_rows - 1) / band_rows;
    std::atomic<int> next(0), err(0);
    std::mutex emu;
    std::string emsg;
    ndt_b200_stats acc;
    memset(&acc, 0, sizeof acc);
    auto body = [&](Worker *w) {
        cudaSetDevice(w->device);
        int r = 0;
        bool uploaded = false;
        ndt_b200_stats mine;
        memset(&mine, 0, sizeof mine);
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= nbands || err.load()) break;
            if (!uploaded) { r = ndt_b200_upload(w->ctx, fs); uploaded = true; }
            const int y0 = b * band_rows, rows = (H - y0) < band_rows ? (H - y0) : band_rows;
            const size_t off = (size_t)y0 * W;
            ndt_b200_stats st;
            if (!r) r = ndt_b200_render_tile(w->ctx, 0, y0, W, rows, rgba_f64 ? rgba_f64 + 4 * off : NULL,
                                             rgba_u8 ? rgba_u8 + 4 * off : NULL, hit ? hit + off : NULL,
                                             obj_id ? obj_id + off : NULL, inv_depth ? inv_depth + off : NULL, &st);
            if (r) {
                std::lock_guard<std::mutex> lk(emu);
                if (!err.load()) { err.store(r); emsg = ndt_b200_last_error(); }
                break;
            }
            stats_sum(&mine, &st);
        }
        std::lock_guard<std::mutex> lk(emu);
        stats_sum(&acc, &mine);
    };
    /* the persistent workers serve the streaming queue; a frame split runs on short-lived threads over the same
     * contexts, so the two entry points must not be mixed while frames are in flight */
    {
        std::unique_lock<std::mutex> lk(m->mu);
        if (m->in_flight) return ndt_set_error(NDT_B200_E_STATE, "ndt_b200_mgpu_render_frame while streamed frames are in flight");
    }
    std::vector<std::thread> th;
    for (int i = 1; i < nctx && i < nbands; ++i) th.emplace_back(body, &m->w[(size_t)i]);
    body(&m->w[0]);
    for (auto &t : th) t.join();
    if (err.load()) return ndt_set_error(err.load(), "%s", emsg.c_str());
    if (stats) *stats = acc;
    return 0;
}

/* an animation, frame by frame as the host produces the scenes */
extern "C" int ndt_b200_mgpu_submit(ndt_b200_mgpu *m, const ndt_flat_scene *fs, uint8_t *rgba_u8, double *rgba_f64)
{
    if (!m || !fs) return ndt_set_error(NDT_B200_E_ARG, "ndt_b200_mgpu_submit: NULL argument");
    int r = ndt_b200_flat_validate(fs, (size_t)fs->h.total_bytes);
    if (r) return r;
    ndt_flat_scene *copy = (ndt_flat_scene *)malloc((size_t)fs->h.total_bytes);
    if (!copy) return ndt_set_error(NDT_B200_E_NOMEM, "out of memory");
    memcpy(copy, fs, (size_t)fs->h.total_bytes);
    {
        std::unique_lock<std::mutex> lk(m->mu);
        /* bounded: the producer (scene_setup + kd build + flatten) does not run away from the GPUs */
        m->cv_done.wait(lk, [&] { return m->in_flight < m->max_queue || m->err; });
        if (m->err) { free(copy); return ndt_set_error(m->err, "%s", m->err_msg.c_str()); }
        Job j = { copy, rgba_u8, rgba_f64 };
        m->jobs.push_back(j);
        ++m->in_flight;
    }
    m->cv_job.notify_one();
    return 0;
}

extern "C" int ndt_b200_mgpu_wait(ndt_b200_mgpu *m, ndt_b200_stats *stats)
{
    if (!m) return ndt_set_error(NDT_B200_E_ARG, "NULL mgpu");
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->in_flight == 0; });
    const int e = m->err;
    if (stats) *stats = m->acc;
    memset(&m->acc, 0, sizeof m->acc);
    if (e) {
        m->err = 0;
        return ndt_set_error(e, "%s", m->err_msg.c_str());
    }
    return 0;
}

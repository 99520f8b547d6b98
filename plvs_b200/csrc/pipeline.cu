// Stream driver: the reference's thread structure around the three replaced surfaces, in its own host language.
// PLVS runs frame construction + Tracking in the caller's thread, LocalMapping and PointCloudMapping in threads of their own
// (src/System.cc:317-398); here the four roles are four std::threads connected by queues, each with its own library handle (= CUDA stream):
//   frame construction : ORBextractor::operator() on batches of frames        (src/Frame.cc:292-330 -> src/ORBextractor.cc:1245); n_ex / 2 threads, alternate batches
//   Tracking           : SearchByProjection(Cur, Last) + SearchByProjection(F, local map points) per frame   (src/Tracking.cc:3593, 4477)
//   LocalMapping       : SearchForTriangulation against the previous frame    (src/LocalMapping.cc:537)
//   PointCloudMapping  : ChiselServer::IntegrateLastDepthImage per frame       (src/PointCloudMapChisel.cc:100-131)
// Host code only: it calls nothing but the public entry points of include/plvs_b200.h, every call stays synchronous for its caller exactly as
// the reference's functions are, and the overlap is between calls of different stages (batch k+1 extracts while batch k is matched).
// bench.py times this driver; plvs_b200/pipeline.py::HotPath.run_stream is the same pipeline on Python threads.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

using namespace plvs;

namespace {

template <class T>
struct Channel {
    std::mutex mu; std::condition_variable cv; std::deque<T> q;
    void put(const T& v) { { std::lock_guard<std::mutex> l(mu); q.push_back(v); } cv.notify_one(); }
    T get() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return !q.empty(); }); T v = q.front(); q.pop_front(); return v; }
};

struct Workspace {           // what one extractor call hands to the next stages (the caller's vectors in the reference)
    std::vector<plvs_keypoint> kps; std::vector<uint8_t> desc; std::vector<int> n, mono;
};

constexpr int kMaxExtractors = 8;

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" int plvs_pipeline_run(plvs_orb* const* ex, int n_ex, plvs_match* m_track, plvs_match* m_tri, plvs_tsdf* tsdf, const plvs_pipeline_job* job,
                                 plvs_pipeline_stats* out)
{
    if (!ex || n_ex < 2 || n_ex > kMaxExtractors || (n_ex & 1) || !m_track || !m_tri || !tsdf || !job || !out) { set_error("null / bad argument"); return PLVS_EINVAL; }
    for (int i = 0; i < n_ex; ++i) if (!ex[i]) { set_error("null extractor handle"); return PLVS_EINVAL; }
    if (job->batch < 1 || job->n_steps < 0 || job->cap < 1 || !job->frames || !job->gray || !job->depth || !job->poses) { set_error("bad pipeline job"); return PLVS_EINVAL; }
    const int B = job->batch, W = job->width, H = job->height, cap = job->cap;
    const size_t px = (size_t)W * H;
    // Extractor handle / workspace i serves the steps s with s % n_ex == i.  n_ex / 2 frame-construction threads take alternate steps, so two
    // batches can be in extraction at once (a batch is a chain of fourteen mostly narrow kernels) while Tracking and LocalMapping read a third one.
    const int NX = n_ex, n_ex_threads = n_ex / 2;
    std::vector<Workspace> ws(NX);
    for (Workspace& w : ws) { w.kps.resize((size_t)B * cap); w.desc.resize((size_t)B * cap * 32); w.n.assign(B, 0); w.mono.assign(B, 0); }
    std::mutex sync_mu; std::condition_variable sync_cv;       // guards free_slot[], extracted[]: who may write a workspace, which steps are ready
    std::vector<char> free_slot(NX, 1);                        // workspace i may be overwritten by the extractor
    std::vector<char> extracted((size_t)std::max(job->n_steps, 1), 0);
    std::atomic<int> users[kMaxExtractors];                    // consumers (Tracking, LocalMapping) still reading workspace i
    for (int i = 0; i < kMaxExtractors; ++i) users[i] = 0;
    Channel<int> q_tri;                    // step index, -1 = end of stream
    std::atomic<int> failed{0};
    std::mutex err_mu; std::string err_msg;
    std::atomic<long long> keypoints{0}, matches{0};
    double busy[4] = {0, 0, 0, 0};
    auto fail = [&](int rc, const char* what) {
        if (!failed.exchange(rc ? rc : PLVS_EINVAL)) { std::lock_guard<std::mutex> l(err_mu); err_msg = std::string(what) + ": " + plvs_last_error(); }
        { std::lock_guard<std::mutex> l(sync_mu); }
        sync_cv.notify_all();
    };
    auto release_slot = [&](int i) { { std::lock_guard<std::mutex> l(sync_mu); free_slot[i] = 1; } sync_cv.notify_all(); };
    cudaStream_t flush_stream = nullptr;
    if (job->flush_buf && job->flush_bytes) { cudaSetDevice(job->device); cudaStreamCreateWithFlags(&flush_stream, cudaStreamNonBlocking); }

    // ---- one step of each stage ------------------------------------------------------------------------------------------------
    auto extract_step = [&](int s) -> bool {
        const int i = s % NX;
        if (flush_stream) {                                   // evict the previous step's working set from L2 (inside the timed region)
            cudaSetDevice(job->device);
            cudaMemsetAsync(job->flush_buf, s & 0xff, job->flush_bytes, flush_stream);
        }
        const size_t f0 = (size_t)job->first_frame + (size_t)s * B;
        const int rc = plvs_orb_extract_batch(ex[i], B, job->gray + f0 * px, W, H, W, px, job->inputs_on_device, 0, 0, ws[i].kps.data(), ws[i].desc.data(), cap,
                                              ws[i].n.data(), ws[i].mono.data());
        if (rc) { fail(rc, "plvs_orb_extract_batch"); return false; }
        long long k = 0;
        for (int b = 0; b < B; ++b) k += ws[i].n[b];
        keypoints += k;
        return true;
    };
    std::vector<int32_t> a1, a2, m12; std::vector<uint8_t> claimed;       // a1/a2/claimed: Tracking's thread only; m12: LocalMapping's only
    auto track_step = [&](int s) {
        const int i = s % NX;
        long long nm = 0;
        for (int b = 0; b < B && !failed; ++b) {
            const plvs_pipeline_frame& fr = job->frames[(size_t)job->first_frame + (size_t)s * B + b];
            if (!fr.valid) continue;
            plvs_orb_device_view dv;
            int rc = plvs_orb_device_result(ex[i], b, &dv);
            if (rc) { fail(rc, "plvs_orb_device_result"); break; }
            plvs_frame_view cur = job->view_template;      // bounds, grid, scale factors, bf of the stream's frames
            cur.n = dv.n; cur.keys = dv.keys; cur.desc = dv.desc; cur.cache_key = dv.cache_key;
            cur.grid_cell_start = dv.grid_cell_start; cur.grid_sorted = dv.grid_sorted;      // built at frame construction when the caller turned it on
            cur.uright = fr.uright; cur.on_device = fr.uright ? (PLVS_VIEW_ON_DEVICE | PLVS_VIEW_URIGHT_ON_HOST) : PLVS_VIEW_ON_DEVICE;
            a1.assign((size_t)std::max(dv.n, 1), -1); a2.assign((size_t)std::max(dv.n, 1), -1); claimed.assign((size_t)std::max(dv.n, 1), 0);
            int n1 = 0, n2 = 0;
            rc = plvs_match_projection_last(m_track, &cur, fr.ql, fr.n_ql, job->th_last, 0, 0, 1, nullptr, a1.data(), &n1);
            if (rc) { fail(rc, "plvs_match_projection_last"); break; }
            for (int k = 0; k < dv.n; ++k) claimed[k] = a1[k] >= 0;      // F.mvpMapPoints[k] is taken: Observations() > 0
            rc = plvs_match_projection_map(m_track, &cur, fr.qm, fr.n_qm, job->th_map, job->nnratio_map, 0, 50.f, claimed.data(), a2.data(), &n2);
            if (rc) { fail(rc, "plvs_match_projection_map"); break; }
            nm += n1 + n2;
        }
        matches += nm;
    };
    auto tri_step = [&](int s) {
        const int i = s % NX;
        long long nm = 0;
        for (int b = 0; b < B && !failed; ++b) {
            const plvs_pipeline_frame& fr = job->frames[(size_t)job->first_frame + (size_t)s * B + b];
            if (!fr.valid) continue;
            plvs_frame_view k1 = job->view_template;      // KeyFrame 1 = the current frame as the extractor returned it to the host
            k1.n = ws[i].n[b]; k1.keys = ws[i].kps.data() + (size_t)b * cap; k1.desc = ws[i].desc.data() + (size_t)b * cap * 32;
            k1.uright = fr.uright; k1.on_device = 0; k1.cache_key = 0;
            m12.assign((size_t)std::max(k1.n, 1), -1);
            int n3 = 0;
            const int rc = plvs_match_triangulation(m_tri, &k1, &fr.last, &fr.fv_cur, &fr.fv_last, fr.has_cur, fr.has_last, fr.F12, fr.ep, 0, 0, 0, m12.data(), &n3);
            if (rc) { fail(rc, "plvs_match_triangulation"); break; }
            nm += n3;
        }
        matches += nm;
    };
    auto map_step = [&](int s) -> bool {
        for (int b = 0; b < B; ++b) {
            const size_t f = (size_t)job->first_frame + (size_t)s * B + b;
            const int rc = plvs_tsdf_integrate_depth(tsdf, job->depth + f * px, W, H, job->bgr ? job->bgr + f * px * 3 : nullptr, W * 3, job->bgr ? 3 : 0,
                                                     job->poses + f * 12, job->bgr ? PLVS_TSDF_SCAN_COLOR : PLVS_TSDF_SCAN, job->inputs_on_device);
            if (rc) { fail(rc, "plvs_tsdf_integrate_depth"); return false; }
        }
        return true;
    };
    auto map_finish = [&] {
        plvs_tsdf_stats st;
        const int rc = plvs_tsdf_last_stats(tsdf, &st);          // waits for the last scan; surfaces a pool-exhaustion error
        if (rc && !failed) fail(rc, "plvs_tsdf_last_stats");
    };

    const double t_begin = now_s();
    const char* serial = std::getenv("PLVS_PIPELINE_SERIAL");
    if (serial && serial[0] == '1') {
        // one thread, stage after stage per step: what the tests run on the CPU execution model (which has a single host thread)
        for (int s = 0; s < job->n_steps && !failed; ++s) {
            double t = now_s(); if (!extract_step(s)) break; busy[0] += now_s() - t;
            t = now_s(); track_step(s); busy[1] += now_s() - t;
            t = now_s(); tri_step(s); busy[2] += now_s() - t;
            t = now_s(); map_step(s); busy[3] += now_s() - t;
        }
        map_finish();
    } else {
        // ---- the stage threads ------------------------------------------------------------------------------------------------------
        double busy_ex[kMaxExtractors] = {0};
        auto extract_stage = [&](int j) {
            for (int s = j; s < job->n_steps && !failed; s += n_ex_threads) {
                const int i = s % NX;
                { std::unique_lock<std::mutex> l(sync_mu); sync_cv.wait(l, [&] { return free_slot[i] || failed; }); free_slot[i] = 0; }
                if (failed) break;
                const double t0 = now_s();
                if (!extract_step(s)) { release_slot(i); break; }
                busy_ex[j] += now_s() - t0;
                users[i] = 2;
                { std::lock_guard<std::mutex> l(sync_mu); extracted[s] = 1; }
                sync_cv.notify_all();
            }
        };
        auto track_stage = [&] {
            for (int s = 0; s < job->n_steps; ++s) {          // Tracking takes the frames in order
                { std::unique_lock<std::mutex> l(sync_mu); sync_cv.wait(l, [&] { return extracted[s] || failed; }); }
                if (failed) break;
                q_tri.put(s);
                const double t0 = now_s();
                track_step(s);
                if (--users[s % NX] == 0) release_slot(s % NX);
                busy[1] += now_s() - t0;
            }
            q_tri.put(-1);
        };
        auto tri_stage = [&] {
            for (;;) {
                const int s = q_tri.get();
                if (s < 0) break;
                const double t0 = now_s();
                tri_step(s);
                if (--users[s % NX] == 0) release_slot(s % NX);
                busy[2] += now_s() - t0;
            }
        };
        auto map_stage = [&] {
            const double t0 = now_s();
            for (int s = 0; s < job->n_steps && !failed; ++s) if (!map_step(s)) break;
            map_finish();
            busy[3] = now_s() - t0;
        };
        std::vector<std::thread> th;
        for (int j = 0; j < n_ex_threads; ++j) th.emplace_back(extract_stage, j);
        th.emplace_back(track_stage); th.emplace_back(tri_stage); th.emplace_back(map_stage);
        for (std::thread& x : th) x.join();
        for (int j = 0; j < n_ex_threads; ++j) busy[0] += busy_ex[j];          // summed over the frame-construction threads
    }
    if (flush_stream) { cudaStreamSynchronize(flush_stream); cudaStreamDestroy(flush_stream); }
    out->wall_s = now_s() - t_begin;
    out->keypoints = keypoints; out->matches = matches;
    out->busy_extract_s = busy[0]; out->busy_track_s = busy[1]; out->busy_tri_s = busy[2]; out->busy_map_s = busy[3];
    if (failed) { set_error("%s", err_msg.c_str()); return failed; }
    return PLVS_OK;
}

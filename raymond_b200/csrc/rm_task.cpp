// rm_task.cpp — render_tiled / TaskHandle / Message over the device-level renderer.
// Reference: src/trace.rs:62-230.  The reference spawns `worker_count` CPU threads that pull
// tiles from a queue and send Tile messages over an mpsc channel; here one driver thread per
// task feeds the GPU and posts the same messages into a single-consumer queue.

#include <cstdlib>
#include <cstring>

#include "rm_internal.hpp"

using namespace rm;

struct rm_task {
    rm_settings settings{};
    rm_gpu_options options{};
    rm_renderer* renderer = nullptr;
    std::thread driver;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<rm_message> messages;
    bool finished = false;          // alive_thread_count == 0   src/trace.rs:74,89
    int status = RM_OK;
    std::string error;
    rm_stats stats{};
    rm_tile_callback callback = nullptr;
    void* callback_user = nullptr;
};

namespace {

bool make_tile(const TileRect& r, size_t sample_count, const rm_vec3* frame_sums, size_t W, rm_tile* out) {
    out->sample_count = sample_count;
    out->width = r.width; out->height = r.height; out->left = r.left; out->top = r.top;
    out->data = (rm_vec3*)malloc(std::max<size_t>(r.width * r.height, 1) * sizeof(rm_vec3));
    if (!out->data) return false;
    for (size_t y = 0; y < r.height; y++)
        memcpy(out->data + y * r.width, frame_sums + (r.top + y) * W + r.left, r.width * sizeof(rm_vec3));
    return true;
}

void post_tiles(rm_task* t, uint32_t kind, size_t sample_count, const rm_vec3* sums) {
    const size_t W = t->settings.camera_settings.backbuffer_width, H = t->settings.camera_settings.backbuffer_height;
    std::vector<TileRect> tiles = tile_layout(W, H, t->settings.tile_size[0], t->settings.tile_size[1]);
    const int world = t->options.world_size > 1 ? t->options.world_size : 1;
    std::deque<rm_message> batch;
    for (size_t i = 0; i < tiles.size(); i++) {
        if (t->options.partition == RM_PARTITION_TILES && (int)(i % (size_t)world) != t->options.rank) continue;
        rm_message m{};
        m.kind = kind;
        if (!make_tile(tiles[i], sample_count, sums, W, &m.tile)) continue;
        batch.push_back(m);
    }
    std::lock_guard<std::mutex> lk(t->mu);
    for (rm_message& m : batch) t->messages.push_back(m);
    t->cv.notify_all();
}

void drive(rm_task* t) {
    const rm_settings& s = t->settings;
    const size_t W = s.camera_settings.backbuffer_width, H = s.camera_settings.backbuffer_height;
    const size_t world = t->options.world_size > 1 ? (size_t)t->options.world_size : 1;
    const bool split_samples = world > 1 && t->options.partition == RM_PARTITION_SAMPLES;
    // this rank's passes: global samples first, first + stride, ...
    const size_t first = split_samples ? (size_t)t->options.rank : 0;
    const size_t stride = split_samples ? world : 1;
    const size_t total = split_samples ? (s.sample_count > first ? (s.sample_count - first + world - 1) / world : 0) : s.sample_count;
    int st = RM_OK;
    // the frame of running sums lands in a pinned block (cached across tasks): D2H at link speed
    rm_vec3* sums = (rm_vec3*)pinned_acquire(std::max<size_t>(W * H, 1) * sizeof(rm_vec3));
    try {
        if (!sums) throw std::bad_alloc();
        // TileProgressed every `samples_per_iteration` passes (src/trace.rs:217-219)
        const size_t chunk = s.samples_per_iteration ? s.samples_per_iteration : (total ? total : 1);
        size_t done = 0;
        while (done < total && st == RM_OK) {
            const size_t n = std::min(chunk, total - done);
            st = rm_renderer_render(t->renderer, first + done * stride, n, stride);
            done += n;
            if (st == RM_OK && done < total) {
                st = rm_renderer_read_sums(t->renderer, sums);
                if (st == RM_OK) post_tiles(t, RM_TILE_PROGRESSED, done, sums);
            }
        }
        if (st == RM_OK) st = rm_renderer_read_sums(t->renderer, sums);
        if (st == RM_OK) post_tiles(t, RM_TILE_FINISHED, total, sums);      // src/trace.rs:211-212
    } catch (const std::bad_alloc&) {
        st = fail(RM_ERR_OUT_OF_MEMORY, "out of host memory in the render driver");
    }
    pinned_release(sums);
    rm_stats stats{};
    rm_renderer_stats(t->renderer, &stats);
    std::lock_guard<std::mutex> lk(t->mu);
    t->stats = stats;
    t->status = st;
    if (st != RM_OK) t->error = rm_last_error();
    t->finished = true;
    t->cv.notify_all();
}

}  // namespace

extern "C" {

rm_task* rm_render_tiled(const rm_scene* scene, const rm_settings* settings, const rm_gpu_options* options) {
    if (!scene || !settings) { fail(RM_ERR_INVALID_ARGUMENT, "rm_render_tiled: null argument"); return nullptr; }
    rm_task* t = new rm_task();
    t->settings = *settings;
    if (options) t->options = *options;
    // scene upload happens here, on the caller's thread, so a bad scene or a missing GPU is
    // reported synchronously; the device copy is the snapshot (the caller may destroy `scene`)
    t->renderer = rm_renderer_create(scene, settings, options);
    if (!t->renderer) { delete t; return nullptr; }
    t->driver = std::thread(drive, t);
    return t;
}

int rm_task_poll(rm_task* t, rm_message* out) {
    if (!t || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_poll: null argument");
    std::lock_guard<std::mutex> lk(t->mu);
    if (t->messages.empty()) return 0;
    *out = t->messages.front();
    t->messages.pop_front();
    return 1;
}

int rm_task_await(rm_task* t, rm_vec3* out) {
    if (!t || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_await: null argument");
    std::unique_lock<std::mutex> lk(t->mu);
    t->cv.wait(lk, [&] { return t->finished; });
    if (t->status != RM_OK) return fail(t->status, t->error);
    const size_t W = t->settings.camera_settings.backbuffer_width, H = t->settings.camera_settings.backbuffer_height;
    for (size_t i = 0; i < W * H; i++) out[i] = rm_vec3{0.0, 0.0, 0.0};
    std::deque<rm_message> keep;
    while (!t->messages.empty()) {
        rm_message m = t->messages.front();
        t->messages.pop_front();
        if (m.kind != RM_TILE_FINISHED) { rm_tile_free(&m.tile); continue; }   // skipped, not a stop (the reference breaks here)
        const rm_tile& tile = m.tile;
        const double c = (double)tile.sample_count;
        for (size_t y = 0; y < tile.height; y++)
            for (size_t x = 0; x < tile.width; x++) {
                const rm_vec3& v = tile.data[x + y * tile.width];
                out[x + tile.left + (y + tile.top) * W] = rm_vec3{v.x / c, v.y / c, v.z / c};   // src/trace.rs:95-97
            }
        rm_tile_free(&m.tile);
    }
    return RM_OK;
}

int rm_task_set_callback(rm_task* t, rm_tile_callback callback, void* user) {
    if (!t) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_set_callback: null task");
    std::lock_guard<std::mutex> lk(t->mu);
    t->callback = callback;
    t->callback_user = user;
    return RM_OK;
}

// async_await: hand every queued TileProgressed to the callback; stop at anything else   src/trace.rs:119-134
int rm_task_pump(rm_task* t) {
    if (!t) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_pump: null task");
    int delivered = 0;
    for (;;) {
        rm_message m;
        rm_tile_callback cb;
        void* user;
        {
            std::lock_guard<std::mutex> lk(t->mu);
            if (t->messages.empty() || t->messages.front().kind != RM_TILE_PROGRESSED) break;
            m = t->messages.front();
            t->messages.pop_front();
            cb = t->callback;
            user = t->callback_user;
        }
        if (cb) cb(&m.tile, user);
        rm_tile_free(&m.tile);
        delivered++;
    }
    return delivered;
}

int rm_task_finished(rm_task* t) {
    if (!t) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_finished: null task");
    std::lock_guard<std::mutex> lk(t->mu);
    return t->finished ? 1 : 0;
}

int rm_task_stats(rm_task* t, rm_stats* out) {
    if (!t || !out) return fail(RM_ERR_INVALID_ARGUMENT, "rm_task_stats: null argument");
    std::unique_lock<std::mutex> lk(t->mu);
    t->cv.wait(lk, [&] { return t->finished; });
    *out = t->stats;
    return t->status;
}

void rm_task_destroy(rm_task* t) {
    if (!t) return;
    if (t->driver.joinable()) t->driver.join();
    for (rm_message& m : t->messages) rm_tile_free(&m.tile);
    rm_renderer_destroy(t->renderer);
    delete t;
}

}  // extern "C"

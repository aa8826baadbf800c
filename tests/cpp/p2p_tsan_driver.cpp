// Native driver of the direct halo push (mvd_p2p_*) for ThreadSanitizer: the ranks are std::threads calling the C ABI of the
// kernel emulator directly -- no Python, hence no interpreter lock whose hand-overs would order the threads behind
// ThreadSanitizer's back.  Compiled together with csrc/spim_b200.cu (-DSPIM_HOST_EMU -fsanitize=thread) by
// tests/test_tsan_push_protocol.py, once as is and once with -DSPIM_EMU_RELAXED_FLAGS (negative control: with relaxed epoch
// flags the pushed halo data are no longer ordered before their consumers, and ThreadSanitizer has to say so).
// Setup (average, export records, connection) is separated by ordinary barriers; the iteration loop has NO synchronisation
// other than the push / wait kernels themselves.
//   usage: p2p_tsan_driver <gz> <gy> <gx> <iterations>
#include "../../include/spim_mvdecon.h"
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct Barrier {
    std::mutex m; std::condition_variable cv; int n, count = 0, gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> l(m);
        const int g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};

static int g_fail = 0;
#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "rank %d: %s failed: %s\n", r, #x, mvd_last_error()); __atomic_store_n(&g_fail, 1, __ATOMIC_RELAXED); } } while (0)

int main(int argc, char** argv) {
    const int grid[3] = {argc > 1 ? atoi(argv[1]) : 2, argc > 2 ? atoi(argv[2]) : 2, argc > 3 ? atoi(argv[3]) : 2};
    const int iters = argc > 4 ? atoi(argv[4]) : 3;
    const int world = grid[0] * grid[1] * grid[2];
    const int brick[3] = {6, 7, 8}, V = 2, ks = 5;
    Barrier bar(world);
    std::vector<std::vector<unsigned char>> records(world, std::vector<unsigned char>(MVD_P2P_RECORD_BYTES));
    std::vector<double> partial(world * 6, 0.0);
    std::vector<int> timed_out(world, 0);
    auto rank_main = [&](int r) {
        const int c[3] = {r / (grid[1] * grid[2]), (r / grid[2]) % grid[1], r % grid[2]};
        mvd_params p;
        mvd_params_default(&p);
        for (int d = 0; d < 3; ++d) p.dims[d] = brick[d];
        p.num_views = V; p.iteration_type = MVD_EFFICIENT_BAYESIAN; p.generation = 2; p.haloed = 1;
        mvd_session* s = nullptr;
        CHECK(mvd_session_create(&p, &s));
        const size_t N = (size_t)brick[0] * brick[1] * brick[2];
        std::vector<float> img(N), w(N, 0.5f), psf((size_t)ks * ks * ks);
        unsigned seed = 1234u + 77u * (unsigned)r;
        auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (float)(seed >> 8) * (1.0f / 16777216.0f); };
        for (auto& v : psf) v = 0.1f + rnd();
        const int kd[3] = {ks, ks, ks};
        for (int v = 0; v < V; ++v) {
            for (auto& x : img) x = 0.05f + 0.95f * rnd();
            CHECK(mvd_set_view(s, v, img.data(), w.data(), psf.data(), kd));
        }
        CHECK(mvd_init(s));
        CHECK(mvd_init_partials(s, &partial[(size_t)r * 6]));
        bar.wait();
        double s0 = 0, s1 = 0;
        for (int q = 0; q < world; ++q) { s0 += partial[(size_t)q * 6]; s1 += partial[(size_t)q * 6 + 1]; }
        CHECK(mvd_set_avg(s, s1 > 0 ? s0 / s1 : 0.5, 1.0));
        mvd_info info;
        CHECK(mvd_get_info(s, &info));
        void* ptr = nullptr; int dims[3], origin[3];
        CHECK(mvd_get_device_buffer(s, 0, &ptr, dims, origin));
        CHECK(mvd_p2p_export(s, records[r].data()));
        bar.wait();
        // neighbour pieces exactly as bricks.BrickRunner._plan_exchange / _boxes derive them
        std::vector<unsigned char> recs; std::vector<int> boxes, slots;
        int lo_mask = 0, hi_mask = 0, npieces = 0;
        for (int d = 0; d < 3; ++d) { if (c[d] > 0) lo_mask |= 1 << d; if (c[d] + 1 < grid[d]) hi_mask |= 1 << d; }
        auto box = [&](const int off[3], int send[6], int recv[6]) -> bool {
            for (int d = 0; d < 3; ++d) {
                const int o = origin[d], n = brick[d], wlo = info.halo_lo[d], whi = info.halo_hi[d];
                if (off[d] == 0) { send[d] = o; send[3 + d] = n; recv[d] = o; recv[3 + d] = n; }
                else if (off[d] < 0) { if (!wlo || !whi) return false; send[d] = o; send[3 + d] = whi; recv[d] = o - wlo; recv[3 + d] = wlo; }
                else { if (!wlo || !whi) return false; send[d] = o + n - wlo; send[3 + d] = wlo; recv[d] = o + n; recv[3 + d] = whi; }
            }
            return true;
        };
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            const int off[3] = {dz, dy, dx}, neg[3] = {-dz, -dy, -dx};
            if (!dz && !dy && !dx) continue;
            const int q[3] = {c[0] + dz, c[1] + dy, c[2] + dx};
            if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= grid[0] || q[1] >= grid[1] || q[2] >= grid[2]) continue;
            int send[6], recv[6], nsend[6], nrecv[6];
            if (!box(off, send, recv) || !box(neg, nsend, nrecv)) continue;
            const int peer = (q[0] * grid[1] + q[1]) * grid[2] + q[2];
            recs.insert(recs.end(), records[peer].begin(), records[peer].end());
            for (int i = 0; i < 6; ++i) boxes.push_back(send[i]);
            for (int i = 0; i < 3; ++i) boxes.push_back(nrecv[i]);
            slots.push_back((neg[0] + 1) * 9 + (neg[1] + 1) * 3 + (neg[2] + 1));
            slots.push_back((off[0] + 1) * 9 + (off[1] + 1) * 3 + (off[2] + 1));
            ++npieces;
        }
        CHECK(mvd_set_halo_mask(s, lo_mask, hi_mask));
        CHECK(mvd_p2p_connect(s, npieces, recs.data(), boxes.data(), slots.data()));
        bar.wait();
        // ---- the iteration loop: nothing but the push / wait kernels orders the ranks from here on ------------------------
        for (int it = 0; it < iters; ++it)
            for (int v = 0; v < V; ++v) {
                CHECK(mvd_p2p_push(s, 0)); CHECK(mvd_p2p_wait(s, 0));
                CHECK(mvd_view_phase(s, v, 0, nullptr));
                CHECK(mvd_p2p_push(s, 1)); CHECK(mvd_p2p_wait(s, 1));
                CHECK(mvd_view_phase(s, v, 1, nullptr));
            }
        CHECK(mvd_p2p_status(s, &timed_out[r]));
        bar.wait();
        CHECK(mvd_p2p_disconnect(s));
        bar.wait();
        mvd_session_destroy(s);
    };
    std::vector<std::thread> ts;
    for (int r = 0; r < world; ++r) ts.emplace_back(rank_main, r);
    for (auto& t : ts) t.join();
    int to = 0;
    for (int v : timed_out) to |= v;
    printf("world %d iterations %d failures %d timeouts %d\n", world, iters, g_fail, to);
    printf(!g_fail && !to ? "P2P_DRIVER_OK\n" : "P2P_DRIVER_FAILED\n");
    return (!g_fail && !to) ? 0 : 1;
}

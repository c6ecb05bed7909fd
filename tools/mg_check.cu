// mg_check.cu -- dev tool: multi-GPU runs of the library from ONE process (one thread per GPU), through the
// C ABI only.  Starts in seconds (no Python, no torchrun), which matters when a W-GPU slot is charged W times.
//
//   mg_check check W [rows cols nnz n prime iters]   run `iters` iterations on W GPUs and on 1 GPU, same matrix
//                                                   (generated on each device by the same hash), compare v, tmp, Av, p
//   mg_check time  W [rows cols nnz n prime iters]   time `iters` iterations on W GPUs (per-phase times of rank 0)
//
// The environment switches of the library (BLK_PIECES, BLK_COLBLOCKS, BLK_RECUR, BLK_P2P, ...) apply as usual, so this
// is also the A/B harness for the multi-GPU experiments of DESIGN.md section 10.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../include/blk_lanczos.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long z)
{
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
}

// heavy-tailed rows (row = rows * u^3: the first rows are dense), uniform columns, values in [1, 100)
__global__ void k_gen(long long nnz, int rows, int cols, int32_t *i, int32_t *j, uint32_t *x)
{
        for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nnz; s += (long long)gridDim.x * blockDim.x) {
                unsigned long long a = mix(2 * s + 1), b = mix(2 * s + 2);
                double u = (double)(a >> 11) * (1.0 / 9007199254740992.0);
                long long r = (long long)((double)rows * u * u * u);
                i[s] = (int32_t)(r < rows ? r : rows - 1);
                j[s] = (int32_t)(b % (unsigned long long)cols);
                x[s] = (uint32_t)(1 + (b >> 40) % 99);
        }
}

struct Job {
        int world = 1, rows = 2000000, cols = 2000000, n = 16, iters = 8;
        long long nnz = 60000000;
        uint32_t prime = 2147483647u;
};

struct RankOut {
        std::vector<uint32_t> v, tmp, Av, p;
        double seconds = 0, ph_ms[BLK_PH_COUNT] = {0};
        int iters = 0, stopped = 0;
        std::string err;
};

static void start_block(std::vector<uint32_t> &v, long long count, uint32_t p)
{
        unsigned long long z = 0x1415926535ull;
        v.resize((size_t)count);
        for (long long t = 0; t < count; t++) {
                z = z * 6364136223846793005ull + 1442695040888963407ull;
                v[(size_t)t] = (uint32_t)((z >> 20) % p);
        }
}

static void run_rank(const Job &job, int rank, int world, int device, const void *nccl_id, bool want_state, bool profile, RankOut *out)
{
        auto bail = [&](const char *what) { out->err = std::string(what) + ": " + blk_last_error(); };
        CK(cudaSetDevice(device));
        int32_t *di, *dj; uint32_t *dx;
        CK(cudaMalloc(&di, sizeof(int32_t) * (size_t)job.nnz)); CK(cudaMalloc(&dj, sizeof(int32_t) * (size_t)job.nnz));
        CK(cudaMalloc(&dx, sizeof(uint32_t) * (size_t)job.nnz));
        k_gen<<<148 * 8, 256>>>(job.nnz, job.rows, job.cols, di, dj, dx);
        CK(cudaDeviceSynchronize());
        blk_params prm;
        memset(&prm, 0, sizeof(prm));
        prm.abi_version = BLK_ABI_VERSION;
        prm.nrows = job.rows; prm.ncols = job.cols; prm.nnz = job.nnz;
        prm.Mi = di; prm.Mj = dj; prm.Mx = dx; prm.coo_on_device = 1;
        prm.n = job.n; prm.prime = job.prime; prm.right_kernel = 0;
        prm.device = device; prm.rank = rank; prm.world = world; prm.nccl_id = world > 1 ? nccl_id : nullptr;
        prm.use_graph = -1;
        blk_ctx *ctx = nullptr;
        if (blk_create(&ctx, &prm)) { bail("blk_create"); return; }
        CK(cudaFree(di)); CK(cudaFree(dj)); CK(cudaFree(dx));
        std::vector<uint32_t> v0;
        start_block(v0, (long long)job.rows * job.n, job.prime);
        if (blk_set_state(ctx, v0.data(), nullptr, 0)) { bail("blk_set_state"); return; }
        int32_t it = 0, st = 0;
        if (blk_iterate(ctx, 2, &it, &st)) { bail("blk_iterate (warm-up)"); return; }          // warm-up (and graph build)
        if (profile) blk_set_profiling(ctx, 1);
        auto t0 = std::chrono::steady_clock::now();
        if (blk_iterate(ctx, job.iters, &it, &st)) { bail("blk_iterate"); return; }
        out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out->iters = it; out->stopped = st;
        if (profile) { int64_t l[BLK_PH_COUNT]; blk_get_phase_times(ctx, out->ph_ms, l); }
        if (want_state) {
                // blk_get_state is collective (the blocks are gathered over NVLink): every rank asks, rank 0's copy is compared
                const size_t pad = (size_t)blk_block_pad(job.rows, job.cols, job.n, 0);
                out->v.assign(pad, 0); out->tmp.assign(pad, 0); out->Av.assign(pad, 0); out->p.assign(pad, 0);
                if (blk_get_state(ctx, out->v.data(), out->tmp.data(), out->Av.data(), out->p.data())) { bail("blk_get_state"); return; }
                if (rank != 0) { std::vector<uint32_t>().swap(out->v); std::vector<uint32_t>().swap(out->tmp); std::vector<uint32_t>().swap(out->Av); std::vector<uint32_t>().swap(out->p); }
        }
        blk_destroy(ctx);
}

static bool run_world(const Job &job, int world, bool want_state, bool profile, std::vector<RankOut> *outs)
{
        unsigned char id[BLK_NCCL_ID_BYTES] = {0};
        if (world > 1 && blk_nccl_unique_id(id)) { printf("FAIL blk_nccl_unique_id: %s\n", blk_last_error()); return false; }
        outs->assign((size_t)world, RankOut());
        std::vector<std::thread> th;
        for (int r = 0; r < world; r++)
                th.emplace_back(run_rank, std::cref(job), r, world, r, (const void *)id, want_state, profile && r == 0, &(*outs)[(size_t)r]);
        for (auto &t : th) t.join();
        bool ok = true;
        for (int r = 0; r < world; r++)
                if (!(*outs)[(size_t)r].err.empty()) { printf("FAIL rank %d: %s\n", r, (*outs)[(size_t)r].err.c_str()); ok = false; }
        return ok;
}

int main(int argc, char **argv)
{
        setvbuf(stdout, nullptr, _IOLBF, 0);          // a run killed by `timeout` keeps what it printed
        if (argc < 3) { printf("usage: %s check|time W [rows cols nnz n prime iters]\n", argv[0]); return 2; }
        Job job;
        const bool check = !strcmp(argv[1], "check");
        job.world = atoi(argv[2]);
        if (argc > 3) job.rows = atoi(argv[3]);
        if (argc > 4) job.cols = atoi(argv[4]);
        if (argc > 5) job.nnz = atoll(argv[5]);
        if (argc > 6) job.n = atoi(argv[6]);
        if (argc > 7) job.prime = (uint32_t)strtoul(argv[7], nullptr, 10);
        if (argc > 8) job.iters = atoi(argv[8]);
        int ndev = 0;
        CK(cudaGetDeviceCount(&ndev));
        if (ndev < job.world) { printf("SKIP: %d GPUs visible, %d wanted\n", ndev, job.world); return 0; }
        printf("%d x %d, %lld nnz, n = %d, p = %u, %d iterations, %d GPUs\n", job.rows, job.cols, job.nnz, job.n, job.prime, job.iters, job.world);
        std::vector<RankOut> mg, one;
        if (!run_world(job, job.world, check, !check, &mg)) return 1;
        double worst = 0;
        for (auto &r : mg) worst = r.seconds > worst ? r.seconds : worst;
        printf("%d GPUs: %.2f ms per iteration (%.2f iterations/s, host clock around blk_iterate, slowest rank)\n", job.world,
               worst / job.iters * 1e3, job.iters / worst);
        if (!check) {
                static const char *names[BLK_PH_COUNT] = {"spmv1", "spmv2", "dots", "small", "ortho", "exchange"};
                printf("  rank 0 phases, ms per iteration (in-stream events; profiling mode serialises nothing but adds event records):");
                for (int k = 0; k < BLK_PH_COUNT; k++) printf(" %s %.2f", names[k], mg[0].ph_ms[k] / job.iters);
                printf("\n");
                return 0;
        }
        if (!run_world(job, 1, true, false, &one)) return 1;
        printf("1 GPU : %.2f ms per iteration\n", one[0].seconds / job.iters * 1e3);
        bool same = mg[0].iters == one[0].iters && mg[0].stopped == one[0].stopped;
        const char *nm[4] = {"v", "tmp", "Av", "p"};
        const std::vector<uint32_t> *a[4] = {&mg[0].v, &mg[0].tmp, &mg[0].Av, &mg[0].p}, *b[4] = {&one[0].v, &one[0].tmp, &one[0].Av, &one[0].p};
        for (int k = 0; k < 4; k++) {
                size_t bad = 0, first = 0;
                for (size_t e = 0; e < a[k]->size(); e++)
                        if ((*a[k])[e] != (*b[k])[e]) { if (!bad) first = e; bad++; }
                printf("  %-3s %s", nm[k], bad ? "MISMATCH" : "identical");
                if (bad) printf(" (%zu words, first at row %zu col %zu)", bad, first / job.n, first % job.n);
                printf("\n");
                same = same && !bad;
        }
        printf("mg_check: %s\n", same ? "PASS" : "FAILED");
        return same ? 0 : 1;
}

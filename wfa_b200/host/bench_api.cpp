// Times the reference's call shape on the GPU path (bench.py `e2e.api_value`):
//   results, errs := algn.AlignBatch(qs, ts [][]byte)        (wfa.go:196-201, result pool wfa_cigar.go:62-96)
// through the C++ mirror of the Go API (wfa.hpp): per-pair byte strings in, flattened by the
// mirror into its page-locked pool, one C-ABI call, result objects out.  Synthetic pairs of the
// named workload from the shared generator (libwfagen.so).  Prints one JSON line.
//   bench_api <config number> <L> <edits> <pairs> <global 0/1> <adaptive 0/1> <steps> <warmup> [device]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "wfa.hpp"

extern "C" {
uint64_t wfagen_stride(uint32_t L, uint32_t nedits, uint32_t window);
void wfagen_pairs(uint64_t base_seed, uint64_t first, uint64_t n_pairs, uint32_t L, uint32_t nedits, uint32_t window, uint32_t max_start,
                  uint8_t *out, uint64_t *q_off, uint32_t *q_len, uint64_t *t_off, uint32_t *t_len, int threads);
}

int main(int argc, char **argv)
{
    if (argc < 9) { std::fprintf(stderr, "usage: bench_api <config> <L> <edits> <pairs> <global> <adaptive> <steps> <warmup> [device]\n"); return 64; }
    const uint32_t config = (uint32_t)atoi(argv[1]), L = (uint32_t)atoi(argv[2]), edits = (uint32_t)atoi(argv[3]);
    const uint64_t n = strtoull(argv[4], nullptr, 10);
    const bool global = atoi(argv[5]) != 0, adaptive = atoi(argv[6]) != 0;
    const int steps = atoi(argv[7]), warmup = atoi(argv[8]), device = argc > 9 ? atoi(argv[9]) : 0;
    const uint64_t stride = wfagen_stride(L, edits, 0);
    std::vector<uint8_t> pool(n * stride + 64);
    std::vector<uint64_t> qo(n), to(n); std::vector<uint32_t> ql(n), tl(n);
    wfagen_pairs(0x57464100ull + config, 0, n, L, edits, 0, 0, pool.data(), qo.data(), ql.data(), to.data(), tl.data(), 16);
    std::vector<std::string> qs(n), ts(n);
    for (uint64_t i = 0; i < n; i++) { qs[i].assign((const char *)pool.data() + qo[i], ql[i]); ts[i].assign((const char *)pool.data() + to[i], tl[i]); }
    std::vector<uint8_t>().swap(pool);

    wfa::Penalties p = {4, 6, 2};
    wfa::Options opt = {global};
    wfa::Aligner *algn = wfa::New(&p, &opt, device);
    if (!algn->ok()) { std::fprintf(stderr, "New failed: %s\n", algn->error().c_str()); return 2; }
    wfa::AdaptiveReductionOption ad = {10, 50, 1};
    if (adaptive && algn->AdaptiveReduction(&ad) != nullptr) return 3;
    std::vector<wfa::AlignmentResult *> rs; std::vector<wfa::Error> es;
    double best = 1e30, sum = 0.0, fl = 0.0, cl = 0.0, ob = 0.0; uint64_t checksum = 0;
    for (int it = 0; it < warmup + steps; it++) {
        const auto t0 = std::chrono::steady_clock::now();
        wfa::Error e = algn->AlignBatch(qs, ts, &rs, &es);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (e) { std::fprintf(stderr, "AlignBatch failed: %s\n", e); return 4; }
        if (it >= warmup) { sum += ms; best = ms < best ? ms : best; fl += algn->last_flatten_ms; cl += algn->last_call_ms; ob += algn->last_objects_ms; }
    }
    uint64_t ok = 0;
    for (uint64_t i = 0; i < n; i++) if (rs[i]) { ok++; checksum += rs[i]->Score + rs[i]->Ops.size() * 31u + (rs[i]->Ops.size() ? rs[i]->Ops[0] : 0); }
    std::printf("{\"api_value\": %.1f, \"ms_per_call_mean\": %.4f, \"ms_per_call_min\": %.4f, \"pairs\": %llu, \"pairs_ok\": %llu, \"steps\": %d, \"checksum\": %llu, \"flatten_ms\": %.3f, \"c_abi_calls_ms\": %.3f, \"objects_ms\": %.3f, "
                "\"call\": \"wfa::Aligner::AlignBatch(vector<string>, vector<string>) -> vector<AlignmentResult*> (wfa.hpp, mirror of the Go API)\"}\n",
                (double)n / (sum / steps / 1e3), sum / steps, best, (unsigned long long)n, (unsigned long long)ok, steps, (unsigned long long)checksum, fl / steps, cl / steps, ob / steps);
    wfa::RecycleAligner(algn);
    return 0;
}

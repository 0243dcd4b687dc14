// Micro-probe: where an FPS iteration spends its cycles (clock64 stamps of one thread over 64 iterations).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DFPS_PROBE -DFPS_PROBE_TID=0 \
//        -I../pcd_reg_hregnet_b200/csrc fps_probe.cu -o fps_probe
// stamps: 0 loop top, 1 after update + warp REDUX, 2 after the packet is sent, 3 all packets polled, 4 winner known
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pcd_reg_hregnet_b200/csrc/fps.cu"

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 16384, M = argc > 3 ? atoi(argv[3]) : 1024;
    const bool weighted = argc > 4 && atoi(argv[4]);
    const int directCS = argc > 5 ? atoi(argv[5]) : 0, directLog2T = argc > 6 ? atoi(argv[6]) : 10;   // direct launch
    const int variant = argc > 7 ? atoi(argv[7]) : 0;           // 0: <512,16> one residue, 1: <256,16> two residues per thread
    auto direct = [&](float* xyz, int32_t* idx) -> int {
        if (variant == 1) return launch_fps_cluster<256, 16, false, 2>(xyz, nullptr, nullptr, idx, B, N, M, directLog2T, directCS, 0);
        return launch_fps_cluster<512, 16, false>(xyz, nullptr, nullptr, idx, B, N, M, directLog2T, directCS, 0);
    };
    std::vector<float> h((size_t)B * N * 3), hw((size_t)B * N);
    srand(1);
    for (auto& v : h) v = 100.f * rand() / RAND_MAX;
    for (auto& v : hw) v = 0.5f + 1.f * rand() / RAND_MAX;
    float *xyz, *w, *temp; int32_t* idx;
    cudaMalloc(&xyz, h.size() * 4); cudaMalloc(&w, hw.size() * 4); cudaMalloc(&temp, hw.size() * 4); cudaMalloc(&idx, (size_t)B * M * 4);
    cudaMemcpy(xyz, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        int rc = directCS ? direct(xyz, idx) : hrn_fps(xyz, weighted ? w : nullptr, nullptr, idx, B, N, M, nullptr);
        if (rc) { printf("hrn_fps rc=%d\n", rc); return 1; }
    }
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) {
        if (directCS) direct(xyz, idx);
        else hrn_fps(xyz, weighted ? w : nullptr, nullptr, idx, B, N, M, nullptr);
    }
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("sync failed\n"); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    if (directCS) {     // same picks as the library's own dispatch?
        std::vector<int32_t> a((size_t)B * M), b2((size_t)B * M);
        cudaMemcpy(a.data(), idx, a.size() * 4, cudaMemcpyDeviceToHost);
        hrn_fps(xyz, nullptr, nullptr, idx, B, N, M, nullptr);
        cudaMemcpy(b2.data(), idx, a.size() * 4, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b2[i];
        printf("[variant %d vs dispatch: %zu mismatches] ", variant, bad);
    }
    printf("B=%d N=%d M=%d w=%d CS=%d: %.4f ms, %.1f ns/iter\n", B, N, M, (int)weighted, directCS, ms, ms * 1e6 / (M - 1));
#ifdef FPS_PROBE
    long long p[64 * 8];
    cudaMemcpyFromSymbol(p, g_fps_probe, sizeof(p));
    double acc[5] = {0, 0, 0, 0, 0};
    int n = 0;
    for (int j = 0; j + 1 < 64 && 500 + j + 1 < M; ++j, ++n) {
        acc[0] += p[j * 8 + 1] - p[j * 8 + 0]; acc[1] += p[j * 8 + 2] - p[j * 8 + 1]; acc[2] += p[j * 8 + 3] - p[j * 8 + 2];
        acc[3] += p[j * 8 + 4] - p[j * 8 + 3]; acc[4] += p[(j + 1) * 8 + 0] - p[j * 8 + 0];
    }
    { int act = 0; for (int j = 0; j < n; ++j) act += (int)p[j * 8 + 5]; printf("  culled kernel: this warp's update ran (per stale bound) in %d of %d probed iterations\n", act, n); }
    double a2[4] = {0, 0, 0, 0};
    for (int j = 0; j < n; ++j) { a2[0] += p[j * 8 + 5] - p[j * 8 + 1]; a2[1] += p[j * 8 + 6] - p[j * 8 + 5]; a2[2] += p[j * 8 + 7] - p[j * 8 + 6]; a2[3] += p[j * 8 + 2] - p[j * 8 + 7]; }
    if (n) printf("  resolve detail: hold/pf/LDS %.0f | REDUX.min %.0f | 4x REDUX.OR %.0f | build+store %.0f\n", a2[0] / n, a2[1] / n, a2[2] / n, a2[3] / n);
    if (n) printf("  cycles/iter (tid %d): update+redux %.0f | resolve+send %.0f | poll %.0f | pick %.0f | total %.0f\n", FPS_PROBE_TID,
                  acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n);
#endif
    return 0;
}

// The operand-image kernels of the tcgen05 likelihood path (pyipn_b200/csrc/tc_images.cuh, tc_perm.cuh) run on the HOST
// from their own source (tests/emu/cuda_shim.h): test infrastructure.  Input: the binary problem file of tc_emu; output: the
// per-draw scale / epilogue choice, the W image and the Phi image (digit planes, per-bin constants, tile flags) byte for
// byte as the device would lay them out.  tests/test_emu.py compares them with what `tc_emu images` -- the restatement the
// full-size emulation runs on -- produces for the same problem.
//
//   images_emu <in.bin> <out.bin>
#include "cuda_shim.h"

#include "../../pyipn_b200/csrc/tc_images.cuh"

template <class T>
static bool rd(FILE* f, T* p, size_t n) {
    return std::fread(p, sizeof(T), n, f) == n;
}

int main(int argc, char** argv) {
    if (argc < 3) return 64;
    const int hifi_mode = 0;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t hdr[3];
    double ab[2];
    if (!rd(f, hdr, 3) || !rd(f, ab, 2)) return 2;
    const int64_t T = hdr[0], J = hdr[1], G = hdr[2];
    std::vector<double> time(T), expo(T), omega(J), a(J), b(J), delays(G);
    std::vector<int64_t> counts(T);
    if (!rd(f, time.data(), T) || !rd(f, expo.data(), T) || !rd(f, counts.data(), T) || !rd(f, omega.data(), J) ||
        !rd(f, a.data(), J) || !rd(f, b.data(), J) || !rd(f, delays.data(), G))
        return 2;
    std::fclose(f);
    using namespace ipn;

    // draw_const_kernel (loglik_prep.cu): the two fields the images depend on
    DrawConst dc;
    dc.C = 0.0;
    dc.kap = ab[1] / kLn2;
    const double c = std::log2(ab[0] / ab[1]);
    dc.c_hi = (float)c;
    dc.c_lo = std::isfinite(c) ? (float)(c - (double)dc.c_hi) : 0.0f;

    // run_loglik_grid_tc (loglik_tc.cu): geometry and the launch sequence
    const int Kpad = (int)(((2 * J + tc::MIN_CROWS + 31) / 32) * 32);
    const int n_dt = (int)((G + tc::TM - 1) / tc::TM), n_bt = (int)((T + tc::TN - 1) / tc::TN);
    const uint32_t a_plane = (uint32_t)Kpad * tc::TM, b_plane = (uint32_t)Kpad * tc::TN;
    const uint32_t a_bytes = tc::PW * a_plane, b_bytes = tc::PP * b_plane + tc::CONST_BYTES;
    const int n_pblk = (int)((T + 1023) / 1024);
    std::vector<int> blk_cnt((size_t)(n_pblk + 1) * 4), perm((size_t)T + 1);
    std::vector<double> blk_sq((size_t)n_pblk + 1);
    alignas(8) int flags[4] = {0, 0, 0, 0};
    if (n_pblk > 0) shim::launch(tc_perm_count_kernel, n_pblk, 1024, T, counts.data(), blk_cnt.data(), blk_sq.data());
    shim::launch(tc_perm_kernel, n_pblk > 0 ? n_pblk : 1, 1024, T, counts.data(), n_pblk, blk_cnt.data(), blk_sq.data(), perm.data(),
                 flags, hifi_mode);
    TcDraw td;
    shim::launch(tc_draw_kernel, 1, 128, (int64_t)0, 1, (int)J, Kpad - 2 * (int)J, a.data(), b.data(), &dc, flags, 1, &td);
    std::vector<uint8_t> W((size_t)n_dt * a_bytes + 16, 0xAB), P((size_t)n_bt * b_bytes + 16, 0xAB);
    uint8_t* Wp = W.data() + ((16 - (uintptr_t)W.data() % 16) % 16);  // 16-byte aligned like device allocations
    uint8_t* Pp = P.data() + ((16 - (uintptr_t)P.data() % 16) % 16);
    shim::launch3(tc_wimg_kernel, Dim3{(unsigned)Kpad / 16, (unsigned)n_dt, 1}, 128, (int64_t)0, (int)J, Kpad, omega.data(), a.data(),
                  b.data(), &dc, &td, G, delays.data(), (int64_t)0, n_dt, a_bytes, a_plane, Wp);
    std::vector<PhiTabEntry> phitab(kPhiTabSize);
    shim::launch(tc_phitab_kernel, kPhiTabSize / 256, 256, phitab.data());
    if (n_bt > 0)
        shim::launch3(tc_pimg_kernel, Dim3{(unsigned)((n_bt + PIMG_TILES - 1) / PIMG_TILES), 1, 1}, 256, (int64_t)0, (int)J, Kpad, T, time.data(), expo.data(),
                      counts.data(), perm.data(), flags, omega.data(), &dc, &td, (const PhiTabEntry*)phitab.data(), n_bt, b_bytes,
                      b_plane, Pp);

    FILE* g = std::fopen(argv[2], "wb");
    if (!g) return 2;
    const int ih[4] = {Kpad, n_dt, n_bt, td.hifi};
    const float fh[2] = {td.ks, td.sigma};
    std::fwrite(ih, 4, 4, g);
    std::fwrite(fh, 4, 2, g);
    std::fwrite(Wp, 1, (size_t)n_dt * a_bytes, g);
    std::fwrite(Pp, 1, (size_t)n_bt * b_bytes, g);
    std::fclose(g);
    return 0;
}

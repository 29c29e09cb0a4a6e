// Host side above the C ABI, in C++: a single-rank restatement of the reference driver's per-lap call list
// (code/tristanmainloop.F90:107-344) written against include/tristan_gpu.h exactly the way the `#ifdef GPU` branch of the
// Fortran mainloop would be: one tgpu_* call per reference procedure, same order, the reference's
// error convention (print and stop, particles.F90:362-370).  The reference is Fortran and no Fortran toolchain exists in
// this image, so this file plays the part of its compiled host: it loads a uniform two-species plasma with the reference's
// MINSTD generator (aux.F90:82-134), runs `laps` laps in one of three modes and prints what Diagnostics() prints every lap
// (particle totals, output.F90:355-358) plus print_timers()-style phase times (communications.F90:263-301).
//
//   tristan_mainloop [--laps N] [--mode calls|step|mirror] [--n MX0 MY0 MZ0] [--ppc P] [--order 1|2|3] [--ntimes T]
//                    [--filter 1|2] [--highorder 0|1]
//
//   calls  : every procedure of the lap as its own tgpu_* call (resident state)            -- INTEGRATION.md 4(a)
//   step   : tgpu_step (the same lap with the redundant ghost refreshes removed)           -- INTEGRATION.md 4(b)
//   mirror : tgpu_step_mirror, host arrays in and out every lap (the state lives in this program's arrays)
//
// Build (done by __graft_entry__.build()):  g++ -O2 -std=c++17 host/tristan_mainloop.cpp -Iinclude -L<pkg> -ltristan_gpu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tristan_gpu.h"

static void gpu_check(int ierr, const char *what)
{
    if (ierr != 0) {                                   // the reference prints and stops
        std::fprintf(stderr, "tristan_gpu error %d in %s: %s\n", ierr, what, tgpu_last_error());
        std::exit(1);
    }
}
#define CALL(f, ...) gpu_check(f(__VA_ARGS__), #f)

// aux.F90:82-134: MINSTD, value seed / 2^31 rounded to single precision
static float random_(double &dseed)
{
    dseed = std::fmod(16807.0 * dseed, 2147483647.0);
    return (float)(dseed / 2147483648.0);
}

int main(int argc, char **argv)
{
    int laps = 5, mx0 = 32, my0 = 16, mz0 = 16, order = 2, ntimes = 4, filter = 2, highorder = 0;
    float ppc = 8.f;
    std::string mode = "calls";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](int k) { if (i + k >= argc) { std::fprintf(stderr, "missing value after %s\n", a.c_str()); std::exit(2); } };
        if (a == "--laps") { next(1); laps = std::atoi(argv[++i]); }
        else if (a == "--mode") { next(1); mode = argv[++i]; }
        else if (a == "--n") { next(3); mx0 = std::atoi(argv[++i]); my0 = std::atoi(argv[++i]); mz0 = std::atoi(argv[++i]); }
        else if (a == "--ppc") { next(1); ppc = (float)std::atof(argv[++i]); }
        else if (a == "--order") { next(1); order = std::atoi(argv[++i]); }
        else if (a == "--ntimes") { next(1); ntimes = std::atoi(argv[++i]); }
        else if (a == "--filter") { next(1); filter = std::atoi(argv[++i]); }
        else if (a == "--highorder") { next(1); highorder = std::atoi(argv[++i]); }
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (mode != "calls" && mode != "step" && mode != "mirror") { std::fprintf(stderr, "--mode calls|step|mirror\n"); return 2; }

    // ---- what initialize() leaves in the module globals (initialize.F90:113-175, fields.F90:154-377, particles.F90:219-251)
    tgpu_params gp;
    std::memset(&gp, 0, sizeof gp);
    gp.dim = 3; gp.order = order;
    int32_t geo[6];
    CALL(tgpu_ghost_width, gp.dim, gp.order, &gp.nghost, &gp.nghostz);
    CALL(tgpu_decompose, gp.dim, gp.order, mx0, my0, mz0, 1, 1, 1, 0, geo);
    gp.mx = geo[0]; gp.my = geo[1]; gp.mz = geo[2]; gp.mxcum = geo[3]; gp.mycum = geo[4]; gp.mzcum = geo[5];
    gp.c = 0.45f; gp.corr = 1.025f; gp.ntimes = ntimes; gp.filter_kind = filter;
    gp.periodicx = gp.periodicy = gp.periodicz = 1;
    const float c_omp = 10.f, gamma0 = 0.f, me = 1.f, mi = 1.f;                         // particles.F90:219-235
    gp.qe = -(gp.c / c_omp) * (gp.c / c_omp) * std::sqrt(1.f + gamma0 * gamma0) / ((0.5f * ppc) * (1.f + me / mi));
    gp.qi = -gp.qe;
    gp.qme = gp.qe / (me * std::fabs(gp.qi)); gp.qmi = gp.qi / (mi * std::fabs(gp.qi));
    const int g = gp.nghost / 2, gz = gp.nghostz / 2;
    gp.x1in = (float)(g + 1); gp.x2in = (float)(mx0 + gp.nghost - g);                   // particles.F90:339-344
    gp.y1in = (float)(g + 1); gp.y2in = (float)(my0 + gp.nghost - g);
    gp.z1in = (float)(gz + 1); gp.z2in = (float)(mz0 + gp.nghostz - gz);
    gp.rank = 0; gp.sizex = gp.sizey = gp.sizez = 1;
    const long long ncell = (long long)mx0 * my0 * mz0;
    gp.maxptl = (int32_t)(2.5 * ppc * ncell) + 4096; gp.buffsize = 10000;
    gp.quirks = TGPU_Q_REFERENCE; gp.pusher = 0; gp.device = -1; gp.highorder = highorder;
    const int maxhlf = gp.maxptl / 2;

    // ---- fields and particles of this rank (code/fields.F90:81-88, code/particles.F90:51-55,100)
    const size_t lot = (size_t)gp.mx * gp.my * gp.mz;
    std::vector<float> fld[6];
    for (auto &f : fld) f.assign(lot, 0.f);
    std::vector<tgpu_particle> p((size_t)gp.maxptl);
    double dseed = 123457.0;                                                            // communications.F90:228-229, rank 0
    const int npair = (int)(0.5f * ppc * (float)ncell);
    for (int n = 0; n < npair; n++) {
        // ion and electron at the same position (particles.F90:2735-2873), small thermal spread
        tgpu_particle q{};
        q.x = (float)(g + 1) + random_(dseed) * (float)mx0;
        q.y = (float)(g + 1) + random_(dseed) * (float)my0;
        q.z = (float)(gz + 1) + random_(dseed) * (float)mz0;
        q.ch = 1.f; q.proc = 0; q.splitlev = 1;
        tgpu_particle ion = q, lec = q;
        ion.u = 0.05f * (random_(dseed) - 0.5f); ion.v = 0.05f * (random_(dseed) - 0.5f); ion.w = 0.05f * (random_(dseed) - 0.5f);
        lec.u = 0.05f * (random_(dseed) - 0.5f); lec.v = 0.05f * (random_(dseed) - 0.5f); lec.w = 0.05f * (random_(dseed) - 0.5f);
        ion.ind = 2 * n + 1; lec.ind = 2 * n + 2;
        p[(size_t)n] = ion; p[(size_t)maxhlf + n] = lec;
    }
    int ions = npair, lecs = npair;

    tgpu_ctx *gpu = nullptr;
    CALL(tgpu_init, &gp, &gpu);
    if (mode != "mirror") {
        CALL(tgpu_fields_h2d, gpu, fld[0].data(), fld[1].data(), fld[2].data(), fld[3].data(), fld[4].data(), fld[5].data());
        CALL(tgpu_particles_h2d, gpu, p.data(), ions, lecs);
    }
    std::printf("mode %s: %dx%dx%d cells (+%d ghosts), order %d, %d + %d particles\n", mode.c_str(), mx0, my0, mz0, gp.nghost,
                order, ions, lecs);
    CALL(tgpu_set_option, gpu, "timing", mode == "calls" ? 1 : 0);

    for (int lap = 1; lap <= laps; lap++) {
        if (mode == "calls") {
            // tristanmainloop.F90:117-272, line for line
            CALL(tgpu_pre_bc_b, gpu);                                   // :114 (acts only in an all-open 3D box)
            CALL(tgpu_bc_b1, gpu); CALL(tgpu_bc_e1, gpu);               // :117-118
            CALL(tgpu_advance_b_halfstep, gpu);                         // :119
            CALL(tgpu_bc_b1, gpu);                                      // :122
            CALL(tgpu_move_particles, gpu);                             // :134
            CALL(tgpu_advance_b_halfstep, gpu);                         // :139
            CALL(tgpu_bc_b1, gpu);                                      // :140
            CALL(tgpu_bc_b2, gpu);                                      // :145
            CALL(tgpu_post_bc_b, gpu); CALL(tgpu_pre_bc_e, gpu);        // :155, :157
            CALL(tgpu_advance_e_fullstep, gpu);                         // :159
            CALL(tgpu_bc_e2, gpu);                                      // :164
            CALL(tgpu_post_bc_e, gpu);                                  // :165
            CALL(tgpu_reset_currents, gpu);                             // :171
            CALL(tgpu_bc_e1, gpu); CALL(tgpu_bc_b1, gpu);               // :181-182
            CALL(tgpu_deposit_particles, gpu);                          // :183
            CALL(tgpu_exchange_particles, gpu);                         // :190
            CALL(tgpu_exchange_current, gpu);                           // :203
            CALL(tgpu_apply_filter, gpu);                               // :213-229
            CALL(tgpu_add_current, gpu);                                // :242
            CALL(tgpu_inject_others, gpu);                              // :257
            CALL(tgpu_exchange_particles, gpu);                         // :268
            CALL(tgpu_inject_others, gpu);                              // :272
            if (lap % 10 == 0) CALL(tgpu_reorder_particles, gpu);       // particles.F90:398-400
        } else if (mode == "step") {
            CALL(tgpu_step, gpu, 1);
        } else {
            CALL(tgpu_step_mirror, gpu, fld[0].data(), fld[1].data(), fld[2].data(), fld[3].data(), fld[4].data(), fld[5].data(),
                 p.data(), &ions, &lecs);
        }
        if (mode != "mirror") CALL(tgpu_counts, gpu, &ions, &lecs);     // Diagnostics(): output.F90:355-358
        std::printf("lap %4d  ions %d  lecs %d\n", lap, ions, lecs);
    }

    if (mode == "calls") {
        double ms[TGPU_NPHASE];
        CALL(tgpu_timers, gpu, ms, 0);
        static const char *names[TGPU_NPHASE] = {"fields", "mover", "deposit", "part_exch", "cur_exch", "filter", "sort", "bc"};
        for (int i = 0; i < TGPU_NPHASE; i++) std::printf("  %-10s %9.3f ms per lap\n", names[i], ms[i] / laps);
    }
    // field energy as a one-number summary of the final state (what energy() would print, tristanmainloop.F90:384-494)
    if (mode != "mirror")
        CALL(tgpu_fields_d2h, gpu, fld[0].data(), fld[1].data(), fld[2].data(), fld[3].data(), fld[4].data(), fld[5].data());
    double en = 0;                                     // interior cells only: ghost index m is never refreshed
    for (auto &f : fld)
        for (int k = gz; k < gp.mz - gz - 1; k++)
            for (int j = g; j < gp.my - g - 1; j++)
                for (int i = g; i < gp.mx - g - 1; i++) {
                    const float v = f[(size_t)i + (size_t)gp.mx * ((size_t)j + (size_t)gp.my * (size_t)k)];
                    en += (double)v * v;
                }
    std::printf("kernel launches %lld, field energy %.6e, particles %d\n", (long long)tgpu_launch_count(gpu), en, ions + lecs);
    CALL(tgpu_finalize, gpu);
    return ions + lecs == 2 * npair ? 0 : 3;            // periodic box: nobody may be lost
}

// Shape weights and the Boris/Vay push shared by the particle kernels.
//   weights: code/particles_movedeposit.F90:447-467 (o1), 709-790 (o2), 1035-1128 (o3);
//            code/particles.F90:738-769, 928-982, 1175-1260 (same, with `shift`)
//   push   : code/particles_movedeposit.F90:861-929
#pragma once

// ---- inline-PTX helpers shared by the cell-run kernels (cellrun.cu, cellrun3.cu) ----
// fp32 reduction into global memory, skipped when the addend is exactly zero.  Written as predicated PTX so that the
// compiler emits `@p RED` instead of a branch + reconvergence barrier around every atomic.
__device__ __forceinline__ void red_nz(float *p, float v)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.neu.f32 p, %1, 0f00000000;\n\t@p red.global.add.f32 [%0], %1;\n\t}"
                 :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add(float *p, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// Asynchronous 4-byte global -> shared copy (LDGSTS): the particle record of the NEXT 16-particle step of a half-warp is
// fetched into per-lane landing slots while the current step is being deposited, so neither the permutation lookup nor
// the record loads sit on the critical path of phase 1 and no registers are held across phase 2.
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }


// S has 8 entries; entries 1..6 are the reference's Sx(1:6); slot 3 <-> cell aint(x).
template <int ORDER>
__device__ __forceinline__ void shape_slots(float d, int shift, float S[8], int &smin, int &smax)
{
    const float half = 1.f / 2.f, quart = 1.f / 4.f, one = 1.f, two = 2.f, thhalf = 3.f / 2.f, nineighth = 9.f / 8.f,
                twoth = 2.f / 3.f, sixth = 1.f / 6.f, negsixth = -1.f / 6.f, negone = -1.f;
#pragma unroll
    for (int i = 0; i < 8; i++) S[i] = 0.f;
    if (ORDER <= 1) {
        S[3 + shift] = 1.f - d; S[4 + shift] = d;
        smin = 3 + shift; smax = 4 + shift;
    } else if (ORDER == 2) {
        if (d <= half) {
            float s2 = half * (d * d - d + quart), s4 = s2 + d, s3 = one - s4 - s2;
            S[2 + shift] = s2; S[3 + shift] = s3; S[4 + shift] = s4;
            smin = 2 + shift; smax = 4 + shift;
        } else {
            float s3 = nineighth - thhalf * d + half * d * d, s5 = s3 - one + d, s4 = one - s5 - s3;
            S[3 + shift] = s3; S[4 + shift] = s4; S[5 + shift] = s5;
            smin = 3 + shift; smax = 5 + shift;
        }
    } else {
        float s2, s3, s4, s5;
        if (d <= half) {
            s2 = negsixth * (d - one) * (d - one) * (d - one);
            s3 = twoth + half * (d - two) * d * d;
            s5 = sixth * d * d * d;
            s4 = one - s5 - s3 - s2;
        } else {
            s5 = sixth * d * d * d;
            s4 = twoth + half * (negone - d) * (one - d) * (one - d);
            s2 = sixth * (one - d) * (one - d) * (one - d);
            s3 = one - s5 - s4 - s2;
        }
        S[2 + shift] = s2; S[3 + shift] = s3; S[4 + shift] = s4; S[5 + shift] = s5;
        smin = 2 + shift; smax = 5 + shift;
    }
}

// Window form for the cell-run kernels: the four weights on slots 2..5 (window index 0..3) of the cell `base`,
// for a particle whose own cell is base + shift (|shift| <= 1).  Orders 1 and 2 only: their support never
// leaves slots 2..5 when |dx| < 1/2 cell per step.
template <int ORDER>
__device__ __forceinline__ void shape_window(float d, int shift, float W[4])
{
    const float half = 1.f / 2.f, quart = 1.f / 4.f, one = 1.f, thhalf = 3.f / 2.f, nineighth = 9.f / 8.f;
    float a0, a1, a2, a3;           // unshifted weights on slots 2..5
    if (ORDER <= 1) { a0 = 0.f; a1 = 1.f - d; a2 = d; a3 = 0.f; }
    else if (d <= half) { a0 = half * (d * d - d + quart); a2 = a0 + d; a1 = one - a2 - a0; a3 = 0.f; }
    else { a1 = nineighth - thhalf * d + half * d * d; a3 = a1 - one + d; a2 = one - a3 - a1; a0 = 0.f; }
    if (shift == 0) { W[0] = a0; W[1] = a1; W[2] = a2; W[3] = a3; }
    else if (shift > 0) { W[0] = 0.f; W[1] = a0; W[2] = a1; W[3] = a2; }      // a3 == 0 whenever shift = +1 is reachable
    else { W[0] = a1; W[1] = a2; W[2] = a3; W[3] = 0.f; }                      // a0 == 0 whenever shift = -1 is reachable
}

// FAST (cell-run kernels): reciprocals and reciprocal square roots through the SFU (MUFU.RCP / MUFU.RSQ, <= 2 ulp) instead
// of the IEEE division / square-root sequences; the difference is of the order of the FMA-contraction noise that the
// mover tolerance already covers (tests/test_gpu_parity.py).
template <bool FAST = false>
__device__ __forceinline__ void push_particle(float c, int pusher, float ex0, float ey0, float ez0, float bx0, float by0,
                                              float bz0, float &x, float &y, float &z, float &u, float &v, float &w,
                                              float cinv_host = 0.f)
{
    const float cinv = FAST ? cinv_host : 1.f / c;      // cinv_host = the host's 1.f / c (same rounding)
    float u0, v0, w0, u1, v1, w1, g, f;
    if (FAST && pusher != 1) {
        u0 = c * u + ex0; v0 = c * v + ey0; w0 = c * w + ez0;
        g = c * rsqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
        bx0 = g * bx0; by0 = g * by0; bz0 = g * bz0;
        f = __fdividef(2.f, 1.f + bx0 * bx0 + by0 * by0 + bz0 * bz0);
        u1 = (u0 + v0 * bz0 - w0 * by0) * f;
        v1 = (v0 + w0 * bx0 - u0 * bz0) * f;
        w1 = (w0 + u0 * by0 - v0 * bx0) * f;
        u0 = u0 + v1 * bz0 - w1 * by0 + ex0;
        v0 = v0 + w1 * bx0 - u1 * bz0 + ey0;
        w0 = w0 + u1 * by0 - v1 * bx0 + ez0;
        u = u0 * cinv; v = v0 * cinv; w = w0 * cinv;
        g = c * rsqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
        x = x + u * g * c; y = y + v * g * c; z = z + w * g * c;
        return;
    }
    if (pusher == 1) {
        g = 1.f / sqrtf(1.f + u * u + v * v + w * w);
        float vx0 = c * u * g, vy0 = c * v * g, vz0 = c * w * g;
        u1 = c * u + 2.f * ex0 + vy0 * bz0 - vz0 * by0;
        v1 = c * v + 2.f * ey0 + vz0 * bx0 - vx0 * bz0;
        w1 = c * w + 2.f * ez0 + vx0 * by0 - vy0 * bx0;
        float ustar = cinv * (u1 * bx0 + v1 * by0 + w1 * bz0);
        float sig = cinv * cinv * (c * c + u1 * u1 + v1 * v1 + w1 * w1) - (bx0 * bx0 + by0 * by0 + bz0 * bz0);
        g = 1.f / sqrtf(0.5f * (sig + sqrtf(sig * sig + 4.f * (bx0 * bx0 + by0 * by0 + bz0 * bz0 + ustar * ustar))));
        float tx = bx0 * g, ty = by0 * g, tz = bz0 * g;
        f = 1.f / (1.f + tx * tx + ty * ty + tz * tz);
        u0 = f * (u1 + (u1 * tx + v1 * ty + w1 * tz) * tx + v1 * tz - w1 * ty);
        v0 = f * (v1 + (u1 * tx + v1 * ty + w1 * tz) * ty + w1 * tx - u1 * tz);
        w0 = f * (w1 + (u1 * tx + v1 * ty + w1 * tz) * tz + u1 * ty - v1 * tx);
    } else {
        u0 = c * u + ex0; v0 = c * v + ey0; w0 = c * w + ez0;
        g = c / sqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
        bx0 = g * bx0; by0 = g * by0; bz0 = g * bz0;
        f = 2.f / (1.f + bx0 * bx0 + by0 * by0 + bz0 * bz0);
        u1 = (u0 + v0 * bz0 - w0 * by0) * f;
        v1 = (v0 + w0 * bx0 - u0 * bz0) * f;
        w1 = (w0 + u0 * by0 - v0 * bx0) * f;
        u0 = u0 + v1 * bz0 - w1 * by0 + ex0;
        v0 = v0 + w1 * bx0 - u1 * bz0 + ey0;
        w0 = w0 + u1 * by0 - v1 * bx0 + ez0;
    }
    u = u0 * cinv; v = v0 * cinv; w = w0 * cinv;
    g = c / sqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
    x = x + u * g * c; y = y + v * g * c; z = z + w * g * c;
}

/* Oracle (TEST INFRASTRUCTURE, never linked into the product): plain-C restatement of the reference's
 * explicit viscous-Burgers solver and of its scoring metrics.
 *
 * Follows /root/reference/1D/data/generate_burgers.py:207-299 (burgers_numeric_solve_free; the Cartesian
 * burgers_numeric_solve at :113-205 is the same stencil over the (u0, f) product) and
 * /root/reference/1D/utils/metrics.py:29-34,77-92 (J and |u|>bound exceed masks).
 *
 * Arithmetic contract ("Solver op order"; verified bit-identical to the reference on
 * CPU by tests/golden/solver_*.npz):  all fp32, NO fma contraction (build with -ffp-contract=off),
 *   a  = fp32(1/(2dx)), d = fp32(visc/dx^2), d2 = fp32(-2 visc/dx^2)   (formed in fp64, rounded once)
 *   us = u*u
 *   transport = (-a)*us[i-1] + a*us[i+1]
 *   diffusion = (d*u[i-1] + d2*u[i]) + d*u[i+1]
 *   u[i] <- u[i] + dt*(((-0.5)*transport + diffusion) + f[k][i]),  k = j / record_every,  ghost cells 0.
 * Snapshots after steps record_every, 2*record_every, ...; output row 0 is u0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* u0:[N,s] f:[N,nt,s] out:[N,nt+1,s] */
void oracle_burgers_solve_free(const float* u0, const float* f, float* out, int64_t N, int s, int nt,
                               double visc, double T, double dt_d) {
    const double dx = 1.0 / (double)(s + 1);
    const float a = (float)(1.0 / (2.0 * dx));
    const float na = (float)(-1.0 / (2.0 * dx));
    const float d = (float)(visc * 1.0 / (dx * dx));
    const float d2 = (float)(visc * -2.0 / (dx * dx));
    const float dt = (float)dt_d;
    const int steps = (int)ceil(T / dt_d);
    const int rec = steps / nt;
    float* u = (float*)malloc(sizeof(float) * (size_t)(s + 2));
    float* v = (float*)malloc(sizeof(float) * (size_t)(s + 2));
    float* us = (float*)malloc(sizeof(float) * (size_t)(s + 2));
    for (int64_t n = 0; n < N; ++n) {
        const float* fn = f + n * (int64_t)nt * s;
        float* on = out + n * (int64_t)(nt + 1) * s;
        memcpy(on, u0 + n * (int64_t)s, sizeof(float) * (size_t)s);
        u[0] = 0.f; u[s + 1] = 0.f;
        memcpy(u + 1, u0 + n * (int64_t)s, sizeof(float) * (size_t)s);
        int c = 0;
        for (int j = 0; j < steps; ++j) {
            int k = j / rec;
            if (k >= nt) k = nt - 1;
            const float* fk = fn + (int64_t)k * s;
            for (int i = 0; i < s + 2; ++i) us[i] = u[i] * u[i];
            for (int i = 1; i <= s; ++i) {
                float tr = na * us[i - 1] + a * us[i + 1];
                float di = (d * u[i - 1] + d2 * u[i]) + d * u[i + 1];
                float rhs = (-0.5f * tr + di) + fk[i - 1];
                v[i] = u[i] + dt * rhs;
            }
            v[0] = 0.f; v[s + 1] = 0.f;
            float* tmp = u; u = v; v = tmp;
            if ((j + 1) % rec == 0 && c < nt) {
                memcpy(on + (int64_t)(c + 1) * s, u + 1, sizeof(float) * (size_t)s);
                ++c;
            }
        }
    }
    free(u); free(v); free(us);
}

/* traj:[N,nt1,s] target_final:[N,s] -> J[N] (fp32 mean over x of squared diff of last row, summed
 * sequentially in fp64 then rounded: the tests compare with a tolerance, torch's reduction order is not
 * reproducible), exceed counts: points[N], times[N] (rows with any exceed), sample flag[N]. */
void oracle_burgers_score(const float* traj, const float* target_final, float u_bound, int64_t N, int nt1, int s,
                          float* J, int32_t* pts, int32_t* times, int32_t* flag) {
    for (int64_t n = 0; n < N; ++n) {
        const float* t = traj + n * (int64_t)nt1 * s;
        double acc = 0.0;
        for (int i = 0; i < s; ++i) {
            float dlt = target_final[n * (int64_t)s + i] - t[(int64_t)(nt1 - 1) * s + i];
            acc += (double)(dlt * dlt);
        }
        J[n] = (float)(acc / (double)s);
        int32_t p = 0, tm = 0;
        for (int r = 0; r < nt1; ++r) {
            int any = 0;
            for (int i = 0; i < s; ++i) {
                if (fabsf(t[(int64_t)r * s + i]) > u_bound) { ++p; any = 1; }
            }
            tm += any;
        }
        pts[n] = p; times[n] = tm; flag[n] = tm > 0;
    }
}

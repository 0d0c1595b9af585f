// CPU oracle for the N-body hot path — TEST INFRASTRUCTURE ONLY (see nbody_oracle.h).
//
// Restates /root/reference/samples/nbody.cc (run_step :51-89, query drivers :106-138) and the
// query-3 rules of /root/reference/hw5.cu (:265-309 reach/cost rule, :509-517 selection,
// :545-548 defaults, :598-601 original-index reporting).  Build: see oracle/Makefile
// (g++ -O3 -ffp-contract=off, no -march flags: no FMA contraction, like the reference build
// `g++ -std=c++11 -O3` of samples/Makefile:2,16 on baseline x86-64).
#include "nbody_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// nbody.cc:9-20
namespace param {
const double dt = 60;
const double eps = 1e-3;
const double G = 6.674e-11;
inline double gravity_device_mass(double m0, double t) { return m0 + 0.5 * m0 * fabs(sin(t / 6000)); }
const double planet_radius = 1e7;
const double missile_speed = 1e6;
inline double get_missile_cost(double t) { return 1e5 + 1e3 * t; }
}  // namespace param

// One run_step(step, ...) of nbody.cc:51-89 on planar arrays.  `m` are the *base* masses
// (already zeroed for Q1 / destroyed devices); device modulation is applied here (nbody.cc:61-64).
template <int MODE>
void run_step(int step, int n, double* q, double* v, const double* m, const unsigned char* is_device,
              double* a, double* mj_buf, int nthreads) {
    double* qx = q;
    double* qy = q + n;
    double* qz = q + 2 * n;
    // nbody.cc:61-64 evaluates mj per pair; it depends on (step, j) only, so hoist it (bit-neutral).
    for (int j = 0; j < n; j++) {
        double mj = m[j];
        if (is_device[j]) mj = param::gravity_device_mass(mj, step * param::dt);
        mj_buf[j] = mj;
    }
    // nbody.cc:56-74
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (int i = 0; i < n; i++) {
        double ax = 0, ay = 0, az = 0;
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            double mj = mj_buf[j];
            double dx = qx[j] - qx[i];
            double dy = qy[j] - qy[i];
            double dz = qz[j] - qz[i];
            double r2 = dx * dx + dy * dy + dz * dz + param::eps * param::eps;
            double dist3;
            if (MODE == ORC_MODE_STRICT) {
                dist3 = pow(r2, 1.5);
            } else {
                dist3 = sqrt(r2 * r2 * r2);
            }
            ax += param::G * mj * dx / dist3;
            ay += param::G * mj * dy / dist3;
            az += param::G * mj * dz / dist3;
        }
        a[i] = ax;
        a[i + n] = ay;
        a[i + 2 * n] = az;
    }
    // nbody.cc:77-88
    for (int i = 0; i < 3 * n; i++) v[i] += a[i] * param::dt;
    for (int i = 0; i < 3 * n; i++) q[i] += v[i] * param::dt;
}

void run_step_mode(int mode, int step, int n, double* q, double* v, const double* m,
                   const unsigned char* is_device, double* a, double* mj_buf, int nthreads) {
    if (mode == ORC_MODE_STRICT)
        run_step<ORC_MODE_STRICT>(step, n, q, v, m, is_device, a, mj_buf, nthreads);
    else
        run_step<ORC_MODE_SQRT3>(step, n, q, v, m, is_device, a, mj_buf, nthreads);
}

inline double dist2(int n, const double* q, int i, int j) {
    double dx = q[i] - q[j];
    double dy = q[i + n] - q[j + n];
    double dz = q[i + 2 * n] - q[j + 2 * n];
    return dx * dx + dy * dy + dz * dz;
}

// hw5.cu:273-274 / :303-304
inline bool missile_reached(int n, const double* q, int planet, int d, int step) {
    double missile_dist = (param::missile_speed * param::dt) * step;
    return dist2(n, q, planet, d) < missile_dist * missile_dist;
}

struct Snapshot {
    int step = -2;
    std::vector<double> q, v;
};

// Runs steps step0 .. n_steps from the state (q, v) "at step0".  Fills ev.
// snaps (Q2 only, optional): per device, state at its reach step (hw5.cu:275-284).
void trajectory(int mode, int kind, int n, int planet, int asteroid, double* q, double* v,
                std::vector<double>& m, const unsigned char* is_device, int destroy_device, int step0,
                int n_steps, int nthreads, orc_events* ev, std::vector<Snapshot>* snaps) {
    std::vector<double> a(3 * n), mj(n);
    std::vector<int> devs;
    for (int i = 0; i < n; i++)
        if (is_device[i]) devs.push_back(i);
    if (kind == ORC_KIND_Q1)
        for (int d : devs) m[d] = 0;  // nbody.cc:109-113
    ev->min_d2 = std::numeric_limits<double>::infinity();
    ev->argmin_step = -1;
    ev->hit_step = -2;
    ev->destroyed_step = -2;
    ev->cost = std::numeric_limits<double>::infinity();
    ev->n_reach = (int)devs.size() < 64 ? (int)devs.size() : 64;
    for (int k = 0; k < 64; k++) ev->reach_step[k] = -2;
    ev->steps_done = step0;
    for (int step = step0; step <= n_steps; step++) {
        if (step > step0) run_step_mode(mode, step, n, q, v, m.data(), is_device, a.data(), mj.data(), nthreads);
        ev->steps_done = step;
        double d2 = dist2(n, q, planet, asteroid);
        if (d2 < ev->min_d2) {  // nbody.cc:118-121 (sqrt taken once at the end, hw5.cu:245-247,407)
            ev->min_d2 = d2;
            ev->argmin_step = step;
        }
        if (kind == ORC_KIND_Q1) continue;
        if (kind == ORC_KIND_Q2) {
            // hw5.cu:396 runs before the hit test of the same step (:397)
            for (int k = 0; k < ev->n_reach; k++) {
                if (ev->reach_step[k] == -2 && missile_reached(n, q, planet, devs[k], step)) {
                    ev->reach_step[k] = step;
                    if (snaps) {
                        (*snaps)[k].step = step;
                        (*snaps)[k].q.assign(q, q + 3 * n);
                        (*snaps)[k].v.assign(v, v + 3 * n);
                    }
                }
            }
        }
        // nbody.cc:131-137 / hw5.cu:295-298
        if (d2 < param::planet_radius * param::planet_radius) {
            ev->hit_step = step;
            break;
        }
        if (kind == ORC_KIND_Q3 && ev->destroyed_step == -2 && m[destroy_device] != 0 &&
            missile_reached(n, q, planet, destroy_device, step)) {  // hw5.cu:299-307
            ev->destroyed_step = step;
            ev->cost = param::get_missile_cost((step + 1) * param::dt);
            m[destroy_device] = 0;
        }
    }
}

}  // namespace

extern "C" {

int orc_run_steps(int mode, int n, double* q, double* v, const double* m, const unsigned char* is_device,
                  int step_begin, int step_end, int nthreads) {
    if (n <= 0) return -1;
    std::vector<double> a(3 * n), mj(n);
    for (int step = step_begin + 1; step <= step_end; step++)
        run_step_mode(mode, step, n, q, v, m, is_device, a.data(), mj.data(), nthreads);
    return 0;
}

int orc_trajectory(int mode, int kind, int n, int planet, int asteroid, double* q, double* v,
                   const double* m, const unsigned char* is_device, int destroy_device, int n_steps,
                   int nthreads, orc_events* ev) {
    if (n <= 0 || !ev) return -1;
    std::vector<double> mm(m, m + n);
    trajectory(mode, kind, n, planet, asteroid, q, v, mm, is_device, destroy_device, 0, n_steps, nthreads, ev,
               nullptr);
    return 0;
}

int orc_solve(int mode, int n, int planet, int asteroid, const double* q0, const double* v0, const double* m0,
              const unsigned char* is_device, int n_steps, int nthreads, orc_answer* ans) {
    if (n <= 0 || !ans) return -1;
    std::vector<int> devs;
    for (int i = 0; i < n; i++)
        if (is_device[i]) devs.push_back(i);
    int dc = (int)devs.size();
    if (dc > 64) return -2;
    // small systems: parallelise over trajectories; large ones: over i inside run_step
    bool par_traj = (n < 128) && nthreads > 1;
    int inner = par_traj ? 1 : nthreads;

    orc_events e1, e2;
    std::vector<Snapshot> snaps(dc);
    auto run_q1 = [&]() {
        std::vector<double> q(q0, q0 + 3 * n), v(v0, v0 + 3 * n), m(m0, m0 + n);
        trajectory(mode, ORC_KIND_Q1, n, planet, asteroid, q.data(), v.data(), m, is_device, -1, 0, n_steps, inner,
                   &e1, nullptr);
    };
    auto run_q2 = [&]() {
        std::vector<double> q(q0, q0 + 3 * n), v(v0, v0 + 3 * n), m(m0, m0 + n);
        trajectory(mode, ORC_KIND_Q2, n, planet, asteroid, q.data(), v.data(), m, is_device, -1, 0, n_steps, inner,
                   &e2, &snaps);
    };
    if (par_traj) {
        std::thread t(run_q1);
        run_q2();
        t.join();
    } else {
        run_q1();
        run_q2();
    }
    ans->min_dist = sqrt(e1.min_d2);
    ans->argmin_step = e1.argmin_step;
    ans->hit_time_step = e2.hit_step;
    ans->gravity_device_id = -1;  // hw5.cu:547-548
    ans->missile_cost = 0;
    ans->n_devices = dc;
    for (int k = 0; k < 64; k++) {
        ans->device_index[k] = k < dc ? devs[k] : -1;
        ans->reach_step[k] = k < dc ? e2.reach_step[k] : -2;
        ans->q3_hit_step[k] = -3;
        ans->q3_cost[k] = std::numeric_limits<double>::infinity();
    }
    if (e2.hit_step == -2) return 0;  // hw5.cu:568: query 3 skipped

    std::vector<orc_events> e3(dc);
    auto run_q3 = [&](int k) {
        if (snaps[k].step == -2) return;  // hw5.cu:458
        std::vector<double> m(m0, m0 + n);
        trajectory(mode, ORC_KIND_Q3, n, planet, asteroid, snaps[k].q.data(), snaps[k].v.data(), m, is_device,
                   devs[k], snaps[k].step, n_steps, inner, &e3[k], nullptr);
        ans->q3_hit_step[k] = e3[k].hit_step;
        ans->q3_cost[k] = e3[k].cost;
    };
    if (par_traj) {
        std::vector<std::thread> ts;
        for (int k = 0; k < dc; k++) ts.emplace_back(run_q3, k);
        for (auto& t : ts) t.join();
    } else {
        for (int k = 0; k < dc; k++) run_q3(k);
    }
    // hw5.cu:509-517: cheapest saving device (ties: lowest index, hw5.cu:575-585 sorts by (step, di))
    double best = std::numeric_limits<double>::infinity();
    for (int k = 0; k < dc; k++) {
        if (ans->q3_hit_step[k] == -2 && e3[k].cost < best) {
            best = e3[k].cost;
            ans->gravity_device_id = devs[k];
            ans->missile_cost = best;
        }
    }
    return 0;
}

// nbody.cc:22-39
int orc_read_header(const char* path, int* n, int* planet, int* asteroid) {
    std::ifstream fin(path);
    if (!fin) return -1;
    fin >> *n >> *planet >> *asteroid;
    return fin ? 0 : -1;
}

int orc_read_input(const char* path, int max_n, int* n, int* planet, int* asteroid, double* q, double* v,
                   double* m, unsigned char* is_device) {
    std::ifstream fin(path);
    if (!fin) return -1;
    fin >> *n >> *planet >> *asteroid;
    if (!fin || *n > max_n || *n <= 0) return -2;
    int nn = *n;
    std::string type;
    for (int i = 0; i < nn; i++) {
        fin >> q[i] >> q[i + nn] >> q[i + 2 * nn] >> v[i] >> v[i + nn] >> v[i + 2 * nn] >> m[i] >> type;
        if (!fin) return -3;
        is_device[i] = (type == "device");
    }
    return 0;
}

// nbody.cc:41-49
int orc_write_output(const char* path, double min_dist, int hit_time_step, int gravity_device_id,
                     double missile_cost) {
    std::ofstream fout(path);
    if (!fout) return -1;
    fout << std::scientific << std::setprecision(std::numeric_limits<double>::digits10 + 1) << min_dist << '\n'
         << hit_time_step << '\n'
         << gravity_device_id << ' ' << missile_cost << '\n';
    return fout ? 0 : -1;
}

}  // extern "C"

#ifdef ORACLE_MAIN
// nbody_oracle <input> <output> [n_steps=200000] [mode=0] [threads=all] [kats.json]
int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <input> <output> [n_steps] [mode] [threads] [kats.json]\n", argv[0]);
        return 2;
    }
    int n_steps = argc > 3 ? atoi(argv[3]) : 200000;
    int mode = argc > 4 ? atoi(argv[4]) : ORC_MODE_STRICT;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if (argc > 5 && atoi(argv[5]) > 0) nthreads = atoi(argv[5]);
    int n, planet, asteroid;
    if (orc_read_header(argv[1], &n, &planet, &asteroid)) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 1;
    }
    std::vector<double> q(3 * n), v(3 * n), m(n);
    std::vector<unsigned char> dev(n);
    if (orc_read_input(argv[1], n, &n, &planet, &asteroid, q.data(), v.data(), m.data(), dev.data())) return 1;
    orc_answer ans;
    if (orc_solve(mode, n, planet, asteroid, q.data(), v.data(), m.data(), dev.data(), n_steps, nthreads, &ans))
        return 1;
    if (orc_write_output(argv[2], ans.min_dist, ans.hit_time_step, ans.gravity_device_id, ans.missile_cost))
        return 1;
    if (argc > 6) {
        FILE* f = fopen(argv[6], "w");
        if (!f) return 1;
        fprintf(f, "{\"n\": %d, \"planet\": %d, \"asteroid\": %d, \"n_steps\": %d, \"mode\": %d,\n", n, planet, asteroid,
                n_steps, mode);
        fprintf(f, " \"min_dist\": \"%.16e\", \"argmin_step\": %d, \"hit_time_step\": %d,\n", ans.min_dist,
                ans.argmin_step, ans.hit_time_step);
        fprintf(f, " \"gravity_device_id\": %d, \"missile_cost\": \"%.16e\",\n", ans.gravity_device_id,
                ans.missile_cost);
        fprintf(f, " \"devices\": [");
        for (int k = 0; k < ans.n_devices; k++)
            fprintf(f, "%s{\"index\": %d, \"reach_step\": %d, \"q3_hit_step\": %d}", k ? ", " : "",
                    ans.device_index[k], ans.reach_step[k], ans.q3_hit_step[k]);
        fprintf(f, "]}\n");
        fclose(f);
    }
    return 0;
}
#endif

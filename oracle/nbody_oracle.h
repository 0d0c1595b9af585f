/*
 * CPU oracle for the N-body hot path — TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libnbody_b200.so, hw5) never links or calls it.
 *
 * It restates, on the CPU, the algorithm of the reference:
 *   - samples/nbody.cc:9-20   physics constants, gravity_device_mass, get_missile_cost
 *   - samples/nbody.cc:51-89  run_step (all-pairs accel, v += a*dt, q += v*dt)
 *   - samples/nbody.cc:106-138 query 1 / query 2 drivers
 *   - hw5.cu:265-309, 509-517, 545-548, 568-602  query 3 (the serial sample leaves it as a TODO)
 *
 * Parity is PINNED: tests/test_oracle.py checks it against the reference's goldens
 * (testcases/b*.out, copied as data to tests/golden/testcases/) and, in the build container,
 * against oracle/_ref/nbody (the unmodified samples/nbody.cc compiled in place).
 *
 * Layout: planar arrays q[3*n] = x block, y block, z block (the layout of hw5.cu:93-97).
 */
#ifndef NBODY_ORACLE_H
#define NBODY_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* arithmetic variants of the pair term */
#define ORC_MODE_STRICT 0 /* nbody.cc:65-72 verbatim: pow(r2,1.5), ((G*mj)*d)/dist3, ascending j, no FMA */
#define ORC_MODE_SQRT3 1  /* same, with pow(r2,1.5) replaced by sqrt(r2*r2*r2) (hw5.cu:200-203)           */

/* trajectory kinds */
#define ORC_KIND_Q1 1 /* devices massless for the whole run, track min d2(planet, asteroid)  */
#define ORC_KIND_Q2 2 /* devices active, first hit step, per-device missile reach steps      */
#define ORC_KIND_Q3 3 /* as Q2, device `destroy_device` loses its mass once the missile reaches it */

typedef struct orc_events {
    double min_d2;        /* min over observed steps of d2(planet, asteroid)                 */
    int argmin_step;      /* first step attaining min_d2                                     */
    int hit_step;         /* first step with d2 < planet_radius^2, else -2                   */
    int destroyed_step;   /* Q3: step at which the missile reached the device, else -2       */
    double cost;          /* Q3: 1e5 + 1e3*((destroyed_step+1)*dt), else +inf                */
    int steps_done;       /* last step simulated                                             */
    int n_reach;          /* number of entries used in reach_step                            */
    int reach_step[64];   /* Q2: per device (ascending body index) first missile-reach step, else -2 */
} orc_events;

typedef struct orc_answer {
    double min_dist;
    int hit_time_step;
    int gravity_device_id;
    double missile_cost;
    /* intermediate known answers */
    int argmin_step;
    int n_devices;
    int device_index[64];
    int reach_step[64];
    int q3_hit_step[64]; /* hit step of the trajectory with that device destroyed (-2 = saved, -3 = not simulated) */
    double q3_cost[64];
} orc_answer;

/* Advance steps step_begin+1 .. step_end (each is one nbody.cc run_step(step, ...)). */
int orc_run_steps(int mode, int n, double* q, double* v, const double* m,
                  const unsigned char* is_device, int step_begin, int step_end, int nthreads);

/* One full trajectory from the given (step 0) state; q, v are updated in place. */
int orc_trajectory(int mode, int kind, int n, int planet, int asteroid, double* q, double* v,
                   const double* m, const unsigned char* is_device, int destroy_device,
                   int n_steps, int nthreads, orc_events* ev);

/* The three queries (nbody.cc main + hw5.cu query 3). */
int orc_solve(int mode, int n, int planet, int asteroid, const double* q, const double* v,
              const double* m, const unsigned char* is_device, int n_steps, int nthreads,
              orc_answer* ans);

/* Text formats of nbody.cc:22-49. Returns 0 on success. Arrays must hold max_n bodies. */
int orc_read_header(const char* path, int* n, int* planet, int* asteroid);
int orc_read_input(const char* path, int max_n, int* n, int* planet, int* asteroid, double* q,
                   double* v, double* m, unsigned char* is_device);
int orc_write_output(const char* path, double min_dist, int hit_time_step, int gravity_device_id,
                     double missile_cost);

#ifdef __cplusplus
}
#endif
#endif

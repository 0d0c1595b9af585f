// Text formats of the hw5 assignment and the CLI entry point.
// Input  (nbody.cc:22-39, hw5.cu:86-131): "n planet asteroid" then n lines
//        "qx qy qz vx vy vz m type"; only type == "device" is interpreted (nbody.cc:62).
// Output (nbody.cc:41-49, hw5.cu:133-141): min_dist, hit_time_step, "device_id cost" with doubles
//        in std::scientific at 17 significant digits (== printf "%.16e").
// Unlike hw5.cu:110-130 the bodies are NOT permuted: planet / asteroid / device indexes are passed
// to the kernels, so the reported device id needs no back-map (hw5.cu:601).
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "nb_internal.h"

namespace {

bool slurp(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    fclose(f);
    return true;
}

struct Cursor {
    const char* p;
    bool next_long(long& v) {
        char* e;
        errno = 0;
        v = strtol(p, &e, 10);
        if (e == p) return false;
        p = e;
        return true;
    }
    bool next_double(double& v) {  // strtod: correctly rounded, as operator>> of the reference
        char* e;
        v = strtod(p, &e);
        if (e == p) return false;
        p = e;
        return true;
    }
    bool next_token(const char*& b, size_t& len) {
        while (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') p++;
        if (!*p) return false;
        b = p;
        while (*p && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
        len = (size_t)(p - b);
        return true;
    }
};

}  // namespace

extern "C" {

int nb_read_header(const char* path, int* n, int* planet, int* asteroid) {
    if (!path || !n || !planet || !asteroid) return NB_ERR_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) {
        nb::set_error_detail(std::string("cannot open ") + path);
        return NB_ERR_IO;
    }
    int k = fscanf(f, "%d %d %d", n, planet, asteroid);
    fclose(f);
    if (k != 3 || *n < 1) {
        nb::set_error_detail(std::string("bad header in ") + path);
        return NB_ERR_IO;
    }
    return NB_OK;
}

int nb_read_input(const char* path, int max_n, int* n, int* planet, int* asteroid, double* q, double* v, double* m,
                  unsigned char* is_device) {
    if (!path || !n || !planet || !asteroid || !q || !v || !m || !is_device) return NB_ERR_ARG;
    std::string text;
    if (!slurp(path, text)) {
        nb::set_error_detail(std::string("cannot open ") + path);
        return NB_ERR_IO;
    }
    Cursor c{text.c_str()};
    long a, b, d;
    if (!c.next_long(a) || !c.next_long(b) || !c.next_long(d) || a < 1) {
        nb::set_error_detail(std::string("bad header in ") + path);
        return NB_ERR_IO;
    }
    if (a > max_n) return NB_ERR_ARG;
    if (b < 0 || b >= a || d < 0 || d >= a) {  // the reference indexes its arrays with them unchecked (nbody.cc:118-120)
        nb::set_error_detail(std::string("planet / asteroid index out of range in ") + path);
        return NB_ERR_IO;
    }
    const int nn = (int)a;
    *n = nn, *planet = (int)b, *asteroid = (int)d;
    for (int i = 0; i < nn; i++) {
        const char* tok;
        size_t len;
        if (!c.next_double(q[i]) || !c.next_double(q[i + nn]) || !c.next_double(q[i + 2 * nn]) ||
            !c.next_double(v[i]) || !c.next_double(v[i + nn]) || !c.next_double(v[i + 2 * nn]) ||
            !c.next_double(m[i]) || !c.next_token(tok, len)) {
            nb::set_error_detail(std::string("truncated body line in ") + path);
            return NB_ERR_IO;
        }
        is_device[i] = (len == 6 && memcmp(tok, "device", 6) == 0);
    }
    return NB_OK;
}

int nb_write_output(const char* path, double min_dist, int hit_time_step, int gravity_device_id, double missile_cost) {
    if (!path) return NB_ERR_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) {
        nb::set_error_detail(std::string("cannot write ") + path);
        return NB_ERR_IO;
    }
    int k = fprintf(f, "%.16e\n%d\n%d %.16e\n", min_dist, hit_time_step, gravity_device_id, missile_cost);
    if (fclose(f) != 0 || k < 0) return NB_ERR_IO;
    return NB_OK;
}

// The input format, written: lets a generated or advanced system go through `hw5`, `nbtool` and the reference's own
// samples/nbody.cc alike.  17 significant digits (%.16e) round-trip every double through strtod / operator>>.
int nb_write_input(const char* path, int n, int planet, int asteroid, const double* q, const double* v, const double* m,
                   const unsigned char* is_device) {
    if (!path || n < 1 || !q || !v || !m || !is_device) return NB_ERR_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) {
        nb::set_error_detail(std::string("cannot write ") + path);
        return NB_ERR_IO;
    }
    bool ok = fprintf(f, "%d %d %d\n", n, planet, asteroid) > 0;
    for (int i = 0; i < n && ok; i++) {
        const char* type = is_device[i] ? "device" : i == planet ? "planet" : i == asteroid ? "asteroid" : "star";
        ok = fprintf(f, "%.16e %.16e %.16e %.16e %.16e %.16e %.16e %s\n", q[i], q[i + n], q[i + 2 * n], v[i], v[i + n],
                     v[i + 2 * n], m[i], type) > 0;
    }
    if (fclose(f) != 0 || !ok) return NB_ERR_IO;
    return NB_OK;
}

// Synthetic system of SURVEY.md 8d, config C5: positions uniform in a cube of side 1e13 m centred at
// (-2.0e20, -2.9e20, 1.8e18) (the goldens' neighbourhood), velocities N(0, (1e7 m/s)^2) per component, masses
// log-uniform in [1e20, 1e30] kg, body 0 = planet, body 1 = asteroid, the last n_devices bodies = gravity devices.
// std::mt19937_64 raw output (standardised) with our own transforms (the std:: distributions are not), so the same
// (n, seed) gives the same system wherever libm's log/cos/pow agree.
int nb_generate_system(int n, unsigned long long seed, int n_devices, double* q, double* v, double* m,
                       unsigned char* is_device, int* planet, int* asteroid) {
    if (n < 2 || n_devices < 0 || n_devices > NB_MAX_DEVICES || !q || !v || !m || !is_device || !planet || !asteroid)
        return NB_ERR_ARG;
    std::mt19937_64 rng(seed);
    auto uni = [&]() { return (double)(rng() >> 11) * (1.0 / 9007199254740992.0); };  // [0, 1)
    const double centre[3] = {-2.0e20, -2.9e20, 1.8e18};
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < n; i++) q[(size_t)c * n + i] = centre[c] + (uni() - 0.5) * 1e13;
    for (size_t k = 0; k < 3 * (size_t)n; k += 2) {  // Box-Muller, both variates used
        const double u1 = 1.0 - uni(), u2 = uni();   // u1 in (0, 1]
        const double r = sqrt(-2.0 * log(u1)) * 1e7, a = 6.283185307179586476925286766559 * u2;
        v[k] = r * cos(a);
        if (k + 1 < 3 * (size_t)n) v[k + 1] = r * sin(a);
    }
    for (int i = 0; i < n; i++) m[i] = pow(10.0, 20.0 + 10.0 * uni());
    const int first_dev = n - n_devices < 2 ? 2 : n - n_devices;
    for (int i = 0; i < n; i++) is_device[i] = i >= first_dev ? 1 : 0;
    *planet = 0, *asteroid = 1;
    return NB_OK;
}

// Start-up path: creating a CUDA context costs hundreds of milliseconds per GPU
// and the driver initialises every VISIBLE device, so before the first CUDA call the process narrows
// CUDA_VISIBLE_DEVICES to the GPUs that will get a trajectory (2 + number of devices of the input).
// The GPU count comes from the /dev/nvidiaN nodes (no CUDA or NVML call); a CUDA_VISIBLE_DEVICES set by the user wins.
static int count_gpus_procfs() {
    int cnt = 0;  // device nodes /dev/nvidia0, /dev/nvidia1, ... (containers often hide /proc/driver/nvidia/gpus)
    for (int g = 0; g < 64; g++) {
        char path[32];
        snprintf(path, sizeof path, "/dev/nvidia%d", g);
        if (access(path, F_OK) == 0) cnt++;
    }
    return cnt;
}

// Start-up path of the `hw5` PROCESS only (csrc/hw5.cu calls it before the first CUDA call; the library itself never
// touches the environment): narrows CUDA_VISIBLE_DEVICES to the first `want` GPUs.  Returns the number of GPUs found.
int nb_hw5_narrow_visible_gpus(int want) {
    if (getenv("CUDA_VISIBLE_DEVICES") || want < 1) return -1;  // the user's choice wins
    const int present = count_gpus_procfs();
    if (present > 0 && want < present) {
        std::string list;
        for (int g = 0; g < want; g++) list += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 1);
        if (getenv("NB_VERBOSE")) fprintf(stderr, "nbody_b200: %d GPUs present, using CUDA_VISIBLE_DEVICES=%s\n", present, list.c_str());
    }
    return present;
}

void nb_internal_leak_on_exit(int on);  // nb_host.cu

// hw5 <input> <output> (hw5.cu:532-616)
int nb_hw5_main(const char* input_path, const char* output_path, int n_gpus) {
    const bool verbose = getenv("NB_VERBOSE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (verbose) {
            auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "nbody_b200: %-28s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
            t0 = t1;
        }
    };
    // Start-up path (the reference starts its two GPUs from parallel host threads, hw5.cu:555-567):
    // driver initialisation, context creation, the |sin| table upload and the kernel module load of every GPU that will be
    // used run on helper threads WHILE this thread parses the input; nothing below waits for them before it has to.
    if (n_gpus <= 0) n_gpus = getenv("NB_HW5_GPUS") ? atoi(getenv("NB_HW5_GPUS")) : 1;
    if (n_gpus < 1) n_gpus = 1;
    std::vector<std::thread> warm;
    std::vector<double> warm_s(n_gpus, 0.0);
    for (int g = 0; g < n_gpus; g++)
        warm.emplace_back([g, &warm_s] {
            const auto w0 = std::chrono::steady_clock::now();
            nb_device_warm(g);  // a GPU that does not exist fails here silently and loudly in nb_solve
            warm_s[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
        });
    auto join_warm = [&] {
        for (auto& t : warm)
            if (t.joinable()) t.join();
    };
    int n, planet, asteroid;
    int rc = nb_read_header(input_path, &n, &planet, &asteroid);
    if (rc) {
        join_warm();
        return rc;
    }
    std::vector<double> q(3 * (size_t)n), v(3 * (size_t)n), m(n);
    std::vector<unsigned char> dev(n);
    rc = nb_read_input(input_path, n, &n, &planet, &asteroid, q.data(), v.data(), m.data(), dev.data());
    lap("read input (beside the GPU start-up)");
    join_warm();
    if (rc) return rc;
    if (verbose)
        for (int g = 0; g < n_gpus; g++) fprintf(stderr, "nbody_b200: gpu %d start-up (driver + context + tables + module) %.3f s\n", g, warm_s[g]);
    int n_traj = 2;
    for (int i = 0; i < n; i++) n_traj += dev[i] ? 1 : 0;
    // How many GPUs: the driver initialises every VISIBLE device (0.6-0.9 s each on the B200 boxes, 2.0-2.6 s for four) and
    // a context costs another 0.6-1.6 s per used GPU, while the kernels of b1024 take ~1 s on one GPU and ~0.6 s on two or
    // four (chain / independent plan, nb_host.cu).  Measured end to end on a 2-GPU box: the second context costs more than
    // its kernels save, so the CLI uses ONE GPU unless told otherwise (NB_HW5_GPUS, or the n_gpus argument).
    if (n_gpus > n_traj) n_gpus = n_traj;
    {
        int have = 0;
        rc = nb_device_count(&have);
        if (rc) return rc;
        if (n_gpus > have) n_gpus = have;
    }
    lap("wait for the GPU start-up");
    nb_system sys{n, planet, asteroid, q.data(), v.data(), m.data(), dev.data()};
    nb_answer ans;
    if (getenv("NB_HW5_FAST_EXIT")) nb_internal_leak_on_exit(1);  // set by the hw5 binary: it leaves through _exit
    rc = nb_solve(&sys, nullptr, n_gpus, NB_N_STEPS, NB_MATH_FAST, &ans);
    if (rc) return rc;
    lap("three queries (nb_solve)");
    if (verbose)
        fprintf(stderr, "nbody_b200: %d trajectories on %d GPU(s): gpu %.3f s, solve wall %.3f s, %.3e pairs/s\n",
                ans.n_trajectories, ans.n_gpus_used, ans.gpu_seconds, ans.wall_seconds,
                ans.gpu_seconds > 0 ? ans.pair_interactions / ans.gpu_seconds : 0.0);
    rc = nb_write_output(output_path, ans.min_dist, ans.hit_time_step, ans.gravity_device_id, ans.missile_cost);
    lap("write output");
    return rc;
}

}  // extern "C"

// Text formats of the hw5 assignment and the CLI entry point.
// Input  (nbody.cc:22-39, hw5.cu:86-131): "n planet asteroid" then n lines
//        "qx qy qz vx vy vz m type"; only type == "device" is interpreted (nbody.cc:62).
// Output (nbody.cc:41-49, hw5.cu:133-141): min_dist, hit_time_step, "device_id cost" with doubles
//        in std::scientific at 17 significant digits (== printf "%.16e").
// Unlike hw5.cu:110-130 the bodies are NOT permuted: planet / asteroid / device indexes are passed
// to the kernels, so the reported device id needs no back-map (hw5.cu:601).
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

// hw5 <input> <output> (hw5.cu:532-616)
int nb_hw5_main(const char* input_path, const char* output_path, int n_gpus) {
    int n, planet, asteroid;
    int rc = nb_read_header(input_path, &n, &planet, &asteroid);
    if (rc) return rc;
    std::vector<double> q(3 * (size_t)n), v(3 * (size_t)n), m(n);
    std::vector<unsigned char> dev(n);
    rc = nb_read_input(input_path, n, &n, &planet, &asteroid, q.data(), v.data(), m.data(), dev.data());
    if (rc) return rc;
    if (n_gpus <= 0) {
        rc = nb_device_count(&n_gpus);
        if (rc) return rc;
    }
    nb_system sys{n, planet, asteroid, q.data(), v.data(), m.data(), dev.data()};
    nb_answer ans;
    rc = nb_solve(&sys, nullptr, n_gpus, NB_N_STEPS, NB_MATH_FAST, &ans);
    if (rc) return rc;
    if (getenv("NB_VERBOSE"))
        fprintf(stderr, "nbody_b200: %d trajectories on %d GPU(s): gpu %.3f s, solve wall %.3f s, %.3e pairs/s\n",
                ans.n_trajectories, ans.n_gpus_used, ans.gpu_seconds, ans.wall_seconds,
                ans.gpu_seconds > 0 ? ans.pair_interactions / ans.gpu_seconds : 0.0);
    return nb_write_output(output_path, ans.min_dist, ans.hit_time_step, ans.gravity_device_id, ans.missile_cost);
}

}  // extern "C"

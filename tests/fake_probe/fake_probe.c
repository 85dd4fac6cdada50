/*
 * fake_probe.c — LD_PRELOAD interposer over the PROBE entry points of libb200probe.so (TEST INFRASTRUCTURE).
 *
 * Lets the NATIVE host's active-probe runner (host/cpp/labels.hpp) go through complete rounds on a GPU-less box with scripted
 * probe results, so its policy — busy GPUs skipped, last idle verdict carried over, ENOMEM inconclusive, calibration, cold-cell
 * localisation, gate — can be compared round by round with the Python runner driven by the same script
 * (tests/test_native_plugin.py::test_scripted_probe_rounds_...).  Enumeration, passive NVLink state and everything else still
 * come from the real library over the mock NVML.  A round starts when the runner asks b200probe_device_busy(0).
 *
 * Script = environment, R = round number from 0:
 *   FAKE_R<R>_BUSY  = "1,3"            NVML indices reported busy
 *   FAKE_R<R>_NOMEM = "0"              indices whose HBM sweep returns B200PROBE_ENOMEM
 *   FAKE_R<R>_COPY  = "0:6100,1:1000"  copy GB/s at 1 GiB per index (default 6600; read 7000, write 6900)
 *   FAKE_R<R>_PAIR  = "2>0:620,0>2:621" pair GB/s by NVML index (default = egress 700)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200probe.h"

static int g_round = -1;

static const char* script(const char* what) {
    char name[64];
    snprintf(name, sizeof(name), "FAKE_R%d_%s", g_round < 0 ? 0 : g_round, what);
    return getenv(name);
}
static int in_list(const char* s, int idx) {
    while (s && *s) {
        char* e;
        long v = strtol(s, &e, 10);
        if (e == s) break;
        if (v == idx) return 1;
        s = (*e == ',') ? e + 1 : e;
        if (!*e) break;
    }
    return 0;
}
static double keyed(const char* s, const char* key, double dflt) {     /* "key:value,key:value" */
    size_t n = strlen(key);
    while (s && *s) {
        if (!strncmp(s, key, n) && s[n] == ':') return atof(s + n + 1);
        s = strchr(s, ',');
        if (s) ++s;
    }
    return dflt;
}

int b200probe_device_info(int idx, b200probe_device_t* out) {
    static int (*real)(int, b200probe_device_t*);
    if (!real) real = (int (*)(int, b200probe_device_t*))dlsym(RTLD_NEXT, "b200probe_device_info");
    int rc = real(idx, out);
    if (!rc) out->cuda_ordinal = idx;          /* no CUDA here: pretend ordinal == NVML index */
    return rc;
}

int b200probe_device_busy(int idx, b200probe_busy_t* out) {
    if (idx == 0) ++g_round;
    memset(out, 0, sizeof(*out));
    out->busy = in_list(script("BUSY"), idx);
    out->compute_procs = out->busy;
    return 0;
}

int b200probe_hbm_sweep(int idx, const b200probe_hbm_cfg_t* cfg, b200probe_hbm_result_t* out, int cap, int* n) {
    (void)cfg;
    if (in_list(script("NOMEM"), idx)) return B200PROBE_ENOMEM;
    if (cap < 3) return B200PROBE_ERANGE;
    char key[16];
    snprintf(key, sizeof(key), "%d", idx);
    const double gbs[3] = {7000.0, 6900.0, keyed(script("COPY"), key, 6600.0)};
    const int modes[3] = {B200PROBE_HBM_READ, B200PROBE_HBM_WRITE, B200PROBE_HBM_COPY};
    for (int i = 0; i < 3; ++i) {
        memset(&out[i], 0, sizeof(out[i]));
        out[i].bytes = 1ull << 30; out[i].mode = modes[i]; out[i].gbs_median = out[i].gbs_best = gbs[i]; out[i].verified = 1;
    }
    *n = 3;
    return 0;
}

int b200probe_gemm(int idx, const b200probe_gemm_cfg_t* cfg, b200probe_gemm_result_t* out) {
    (void)idx; (void)cfg;
    memset(out, 0, sizeof(*out));
    out->m = out->n = out->k = 8192; out->tflops_median = 1600.0; out->tflops_best = 1610.0; out->samples = 1024; out->verified = 1;
    return 0;
}

int b200probe_nvlink_a2a(const int* ords, int g, const b200probe_a2a_cfg_t* cfg, double* pair, b200probe_a2a_result_t* out) {
    (void)cfg;
    memset(out, 0, sizeof(*out));
    out->g = g; out->ms_median = out->ms_best = 1.0; out->verified = 1;
    out->pair_source = g > 2 ? B200PROBE_PAIR_STEPPED : B200PROBE_PAIR_ISOLATED;
    double mn = 1e300, mx = 0;
    for (int i = 0; i < g; ++i) {
        out->egress_gbs[i] = out->ingress_gbs[i] = 700.0;
        for (int j = 0; j < g; ++j) {
            double v = 0;
            if (i != j) {
                char key[32];
                snprintf(key, sizeof(key), "%d>%d", ords[i], ords[j]);
                v = keyed(script("PAIR"), key, 700.0);
                if (v < mn) mn = v;
                if (v > mx) mx = v;
            }
            if (pair) pair[i * g + j] = v;
        }
    }
    out->min_pair_gbs = mn; out->max_pair_gbs = mx;
    return 0;
}

int b200probe_hbm_release(int o) { (void)o; return 0; }
int b200probe_gemm_release(int o) { (void)o; return 0; }
int b200probe_a2a_release(void) { return 0; }

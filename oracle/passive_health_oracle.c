/*
 * passive_health_oracle.c — CPU twin of the reference plugin's enumerate + passive health loop.
 * TEST INFRASTRUCTURE ONLY (see oracle.c header for who may load it).
 *
 * PARITY UNPINNED.  north_star names "nvmlDeviceGetHealth"; NVML has no such call (SURVEY.md §0
 * fact 3).  The reference repo selects this behaviour by installing the NVIDIA k8s-device-plugin
 * chart un-pinned (/root/reference/README.md:109,116) with /root/reference/values.yaml:1-18; the
 * plugin's source is not vendored, so the loop below follows
 *   (1) NVML's documented contracts in the local header: nvml.h:4063 nvmlDeviceGetCount_v2,
 *       :4131 GetHandleByIndex_v2, :4616 GetUUID, :4274 GetName, :6110 GetMemoryInfo,
 *       :6161 GetCudaComputeCapability, :9125 EventSetCreate, :9161 RegisterEvents,
 *       :9222 EventSetWait_v2, event constants :2818 :2824 :2835, payload struct :2892;
 *   (2) SURVEY.md §3.3's [RECALLED] description of upstream `checkHealth`:
 *       - DP_DISABLE_HEALTHCHECKS = "all" or containing "xids" disables the loop;
 *       - otherwise its comma list extends the skip set {13,31,43,45,68,109};
 *       - register Xid|DBE|SBE (masked by the supported set) per device; a device that cannot be
 *         queried or registered is Unhealthy from the start;
 *       - wait: TIMEOUT -> nothing; other error -> ALL devices Unhealthy; non-XID event -> ignored;
 *         skipped XID -> ignored; UUID of the event's device unreadable -> ALL Unhealthy;
 *         UUID unknown -> ignored; else that device Unhealthy.  No path back to Healthy.
 * On the GPU box this twin, a pynvml twin (tests/test_gpu_parity.py) and libb200probe.so must
 * agree on device count, UUID order and health strings — the strongest statement available.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <dlfcn.h>
#include <nvml.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORACLE_MAXD 64

typedef struct {
    int index;
    char uuid[96];
    char name[96];
    uint64_t mem_total;
    int cc_major, cc_minor;
    int healthy;    /* 1 = "Healthy", 0 = "Unhealthy" */
} oracle_verdict_t;

static void* L;
static nvmlReturn_t (*pInit)(void);
static nvmlReturn_t (*pShutdown)(void);
static nvmlReturn_t (*pCount)(unsigned int*);
static nvmlReturn_t (*pByIndex)(unsigned int, nvmlDevice_t*);
static nvmlReturn_t (*pUUID)(nvmlDevice_t, char*, unsigned int);
static nvmlReturn_t (*pName)(nvmlDevice_t, char*, unsigned int);
static nvmlReturn_t (*pMem)(nvmlDevice_t, nvmlMemory_t*);
static nvmlReturn_t (*pCC)(nvmlDevice_t, int*, int*);
static nvmlReturn_t (*pSupported)(nvmlDevice_t, unsigned long long*);
static nvmlReturn_t (*pRegister)(nvmlDevice_t, unsigned long long, nvmlEventSet_t);
static nvmlReturn_t (*pSetCreate)(nvmlEventSet_t*);
static nvmlReturn_t (*pSetWait)(nvmlEventSet_t, nvmlEventData_t*, unsigned int);
static nvmlReturn_t (*pSetFree)(nvmlEventSet_t);

static oracle_verdict_t V[ORACLE_MAXD];
static nvmlDevice_t H[ORACLE_MAXD];
static int N;
static nvmlEventSet_t SET;
static int LOOP_DISABLED;
static unsigned long long SKIP[128];
static int NSKIP;

static int load(const char* path) {
    if (!path || !*path) path = "libnvidia-ml.so.1";
    L = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!L) return -1;
#define G(p, n) *(void**)(&p) = dlsym(L, n); if (!p) return -2;
    G(pInit, "nvmlInit_v2") G(pShutdown, "nvmlShutdown") G(pCount, "nvmlDeviceGetCount_v2") G(pByIndex, "nvmlDeviceGetHandleByIndex_v2")
    G(pUUID, "nvmlDeviceGetUUID") G(pName, "nvmlDeviceGetName") G(pMem, "nvmlDeviceGetMemoryInfo") G(pCC, "nvmlDeviceGetCudaComputeCapability")
    G(pSupported, "nvmlDeviceGetSupportedEventTypes") G(pRegister, "nvmlDeviceRegisterEvents") G(pSetCreate, "nvmlEventSetCreate")
    G(pSetWait, "nvmlEventSetWait_v2") G(pSetFree, "nvmlEventSetFree")
#undef G
    return 0;
}

static int enumerate(void) {
    unsigned int c = 0;
    if (pCount(&c) != NVML_SUCCESS) return -3;
    if (c > ORACLE_MAXD) c = ORACLE_MAXD;
    for (unsigned i = 0; i < c; ++i) {
        nvmlMemory_t m;
        memset(&V[i], 0, sizeof(V[i]));
        V[i].index = (int)i;
        V[i].healthy = 1;
        if (pByIndex(i, &H[i]) != NVML_SUCCESS) return -4;
        if (pUUID(H[i], V[i].uuid, sizeof(V[i].uuid)) != NVML_SUCCESS) return -5;
        if (pName(H[i], V[i].name, sizeof(V[i].name)) != NVML_SUCCESS) return -6;
        if (pMem(H[i], &m) != NVML_SUCCESS) return -7;
        V[i].mem_total = m.total;
        if (pCC(H[i], &V[i].cc_major, &V[i].cc_minor) != NVML_SUCCESS) return -8;
    }
    N = (int)c;
    return 0;
}

static void build_skip(const char* spec) {
    static const unsigned long long app[] = {13, 31, 43, 45, 68, 109};
    NSKIP = 0;
    for (unsigned i = 0; i < sizeof(app) / sizeof(app[0]); ++i) SKIP[NSKIP++] = app[i];
    if (!spec) return;
    char* dup = strdup(spec);
    for (char* tok = strtok(dup, ","); tok; tok = strtok(NULL, ",")) {
        while (isspace((unsigned char)*tok)) ++tok;
        char* end = tok + strlen(tok);
        while (end > tok && isspace((unsigned char)end[-1])) *--end = 0;
        if (!*tok) continue;
        int ok = 1;
        for (char* p = tok; *p; ++p) if (!isdigit((unsigned char)*p)) ok = 0;
        if (ok && NSKIP < 128) SKIP[NSKIP++] = strtoull(tok, NULL, 10);
    }
    free(dup);
}

int oracle_ph_open(const char* nvml_path, const char* disable_healthchecks) {
    int rc = load(nvml_path);
    if (rc) return rc;
    if (pInit() != NVML_SUCCESS) return -9;
    rc = enumerate();
    if (rc) return rc;
    char low[256] = "";
    if (disable_healthchecks) {
        size_t i = 0;
        for (; disable_healthchecks[i] && i < sizeof(low) - 1; ++i) low[i] = (char)tolower((unsigned char)disable_healthchecks[i]);
        low[i] = 0;
    }
    LOOP_DISABLED = (!strcmp(low, "all") || strstr(low, "xids")) ? 1 : 0;
    if (LOOP_DISABLED) return 0;
    build_skip(low);
    if (pSetCreate(&SET) != NVML_SUCCESS) return -10;
    const unsigned long long want = nvmlEventTypeXidCriticalError | nvmlEventTypeDoubleBitEccError | nvmlEventTypeSingleBitEccError;
    for (int i = 0; i < N; ++i) {
        unsigned long long sup = 0;
        if (pSupported(H[i], &sup) != NVML_SUCCESS) { V[i].healthy = 0; continue; }
        if (pRegister(H[i], want & sup, SET) != NVML_SUCCESS) V[i].healthy = 0;
    }
    return 0;
}

/* one wait; returns the raw nvmlReturn_t of the wait */
int oracle_ph_poll(int timeout_ms) {
    if (LOOP_DISABLED) return NVML_ERROR_TIMEOUT;
    nvmlEventData_t e;
    memset(&e, 0, sizeof(e));
    nvmlReturn_t r = pSetWait(SET, &e, (unsigned)timeout_ms);
    if (r == NVML_ERROR_TIMEOUT) return r;
    if (r != NVML_SUCCESS) { for (int i = 0; i < N; ++i) V[i].healthy = 0; return r; }
    if (e.eventType != nvmlEventTypeXidCriticalError) return r;
    for (int i = 0; i < NSKIP; ++i) if (SKIP[i] == e.eventData) return r;
    char uuid[96];
    if (pUUID(e.device, uuid, sizeof(uuid)) != NVML_SUCCESS) { for (int i = 0; i < N; ++i) V[i].healthy = 0; return r; }
    for (int i = 0; i < N; ++i) if (!strcmp(V[i].uuid, uuid)) V[i].healthy = 0;
    return r;
}

int oracle_ph_verdicts(oracle_verdict_t* out, int cap, int* n) {
    if (cap < N) return -11;
    memcpy(out, V, sizeof(V[0]) * (size_t)N);
    *n = N;
    return 0;
}

void oracle_ph_close(void) {
    if (!L) return;
    if (SET) pSetFree(SET);
    SET = NULL;
    pShutdown();
    dlclose(L);
    L = NULL;
    N = 0;
}

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

/* CPU baseline: mean microseconds per full enumerate / per zero-timeout poll */
double oracle_ph_time_enumerate(int iters) {
    double t0 = now_us();
    for (int i = 0; i < iters; ++i) if (enumerate()) return -1.0;
    return (now_us() - t0) / iters;
}
double oracle_ph_time_poll(int iters, int timeout_ms) {
    double t0 = now_us();
    for (int i = 0; i < iters; ++i) oracle_ph_poll(timeout_ms);
    return (now_us() - t0) / iters;
}

/* SURVEY.md §8d config 1, N-thread variant: one thread per GPU (up to `threads`), each with its OWN event set registered
 * on its device, `iters` zero-timeout waits after 10 warm-ups.  Returns the mean microseconds per poll seen by a thread;
 * *polls_per_s = aggregate polls per second over all threads (wall clock of the slowest thread). */
typedef struct { int dev, iters; double us; int rc; } poll_thread_t;
static void* poll_thread(void* arg) {
    poll_thread_t* t = (poll_thread_t*)arg;
    nvmlEventSet_t set;
    t->rc = -1;
    if (pSetCreate(&set) != NVML_SUCCESS) return NULL;
    unsigned long long sup = 0;
    const unsigned long long want = nvmlEventTypeXidCriticalError | nvmlEventTypeDoubleBitEccError | nvmlEventTypeSingleBitEccError;
    if (pSupported(H[t->dev], &sup) == NVML_SUCCESS && pRegister(H[t->dev], want & sup, set) == NVML_SUCCESS) {
        nvmlEventData_t e;
        for (int i = 0; i < 10; ++i) pSetWait(set, &e, 0);
        double t0 = now_us();
        for (int i = 0; i < t->iters; ++i) pSetWait(set, &e, 0);
        t->us = (now_us() - t0) / t->iters;
        t->rc = 0;
    }
    pSetFree(set);
    return NULL;
}
double oracle_ph_time_poll_threads(int threads, int iters, double* polls_per_s) {
    if (!L || N <= 0) return -1.0;
    if (threads > N) threads = N;
    if (threads < 1) threads = 1;
    poll_thread_t t[ORACLE_MAXD];
    pthread_t th[ORACLE_MAXD];
    double t0 = now_us();
    for (int i = 0; i < threads; ++i) { t[i].dev = i; t[i].iters = iters; t[i].us = 0; pthread_create(&th[i], NULL, poll_thread, &t[i]); }
    double sum = 0;
    for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); if (t[i].rc) return -2.0; sum += t[i].us; }
    double wall_us = now_us() - t0;
    if (polls_per_s) *polls_per_s = (double)threads * (iters + 10) / (wall_us * 1e-6);
    return sum / threads;
}

/* pin the calling thread to one CPU of its allowed set (the single-thread figures are quoted pinned); returns the CPU or -1 */
int oracle_pin_self(int nth_allowed_cpu) {
    cpu_set_t allowed, one;
    if (sched_getaffinity(0, sizeof(allowed), &allowed)) return -1;
    int seen = 0;
    for (int c = 0; c < CPU_SETSIZE; ++c) {
        if (!CPU_ISSET(c, &allowed)) continue;
        if (seen++ == nth_allowed_cpu) {
            CPU_ZERO(&one);
            CPU_SET(c, &one);
            return pthread_setaffinity_np(pthread_self(), sizeof(one), &one) ? -1 : c;
        }
    }
    return -1;
}

/*
 * mock_nvml.c — a scripted stand-in for libnvidia-ml.so.1 (TEST INFRASTRUCTURE).
 *
 * Lets the enumerate + passive-health path of libb200probe.so and of the oracle twin run on a
 * GPU-less box, and lets tests inject XID/ECC events (SURVEY.md §4 "Fake NVML").  Implements only
 * the NVML entry points the path uses, with the signatures of nvml.h (NVML 12).
 *
 * Environment (read at nvmlInit_v2):
 *   MOCK_NVML_DEVICES=N            number of devices (default 2, max 16)
 *   MOCK_NVML_NO_EVENTS=i,j        nvmlDeviceRegisterEvents -> NOT_SUPPORTED for these
 *   MOCK_NVML_EVENTS_QUERY_FAIL=i  nvmlDeviceGetSupportedEventTypes -> UNKNOWN for these
 *   MOCK_NVML_UUID_FAIL=i          nvmlDeviceGetUUID fails for these (enumeration error path)
 *   MOCK_NVML_LINKS_DOWN=d:l,d:l   link l of device d reports NVML_FEATURE_DISABLED
 *   MOCK_NVML_BUSY=d:procs:util,.. device d reports `procs` foreign compute processes (pids 40000+) and `util` % GPU
 *                                  utilisation (read at every query, so a test can change it between probe rounds)
 *   MOCK_NVML_WAIT_FULL=1          an empty wait blocks for its whole timeout (default: at most 2 ms, to keep tests fast)
 *   MOCK_NVML_EVENT_FILE=path      events for a mock living in ANOTHER process (the native daemon): every
 *                                  wait first queues the lines "kind dev data" appended to the file since
 *                                  the last wait
 * Events are queued with mock_nvml_push(kind, device, data) and popped one per EventSetWait:
 *   kind 0 = XID critical (data = xid), 1 = double-bit ECC, 2 = single-bit ECC,
 *   kind 3 = wait returns error `data` (nvmlReturn_t), kind 4 = XID on a device whose UUID cannot
 *   be read, kind 5 = XID on a device that was never enumerated.
 */
#include <nvml.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXD 16
typedef struct { int idx; int uuid_fail; char uuid[64]; } mdev_t;
static mdev_t g_dev[MAXD + 2];          /* [n] = unreadable-uuid device, [n+1] = never-enumerated device */
static int g_n = 2, g_inited = 0;
static int g_no_events[MAXD], g_query_fail[MAXD];

typedef struct { int kind, dev; unsigned long long data; } mev_t;
static mev_t g_q[256];
static int g_qh = 0, g_qt = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static unsigned long long g_registered[MAXD];
static int g_waits = 0;

static void parse_list(const char* env, int* flags) {
    const char* s = getenv(env);
    memset(flags, 0, sizeof(int) * MAXD);
    if (!s) return;
    while (*s) {
        int v = (int)strtol(s, (char**)&s, 10);
        if (v >= 0 && v < MAXD) flags[v] = 1;
        while (*s == ',' || *s == ' ') ++s;
    }
}

void mock_nvml_push(int kind, int dev, unsigned long long data) {
    pthread_mutex_lock(&g_mu);
    g_q[g_qt % 256] = (mev_t){kind, dev, data};
    g_qt++;
    pthread_mutex_unlock(&g_mu);
}
int mock_nvml_wait_calls(void) { return g_waits; }
unsigned long long mock_nvml_registered(int dev) { return dev >= 0 && dev < MAXD ? g_registered[dev] : 0; }

nvmlReturn_t nvmlInit_v2(void) {
    const char* n = getenv("MOCK_NVML_DEVICES");
    g_n = n ? atoi(n) : 2;
    if (g_n < 0) g_n = 0;
    if (g_n > MAXD) g_n = MAXD;
    int uuid_fail[MAXD];
    parse_list("MOCK_NVML_NO_EVENTS", g_no_events);
    parse_list("MOCK_NVML_EVENTS_QUERY_FAIL", g_query_fail);
    parse_list("MOCK_NVML_UUID_FAIL", uuid_fail);
    for (int i = 0; i < g_n + 2; ++i) {
        g_dev[i].idx = i;
        g_dev[i].uuid_fail = (i < g_n) ? uuid_fail[i] : (i == g_n);
        snprintf(g_dev[i].uuid, sizeof(g_dev[i].uuid), "GPU-b2000000-0000-4000-8000-%012x", i);
    }
    memset(g_registered, 0, sizeof(g_registered));
    g_qh = g_qt = 0;
    g_waits = 0;
    g_inited = 1;
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlShutdown(void) { g_inited = 0; return NVML_SUCCESS; }
const char* nvmlErrorString(nvmlReturn_t r) {
    switch (r) {
        case NVML_SUCCESS: return "Success";
        case NVML_ERROR_TIMEOUT: return "Timeout";
        case NVML_ERROR_NOT_SUPPORTED: return "Not Supported";
        case NVML_ERROR_INVALID_ARGUMENT: return "Invalid Argument";
        case NVML_ERROR_GPU_IS_LOST: return "GPU is lost";
        default: return "Mock NVML error";
    }
}
nvmlReturn_t nvmlSystemGetDriverVersion(char* v, unsigned int len) { snprintf(v, len, "580.00.mock"); return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetCount_v2(unsigned int* c) { if (!g_inited) return NVML_ERROR_UNINITIALIZED; *c = (unsigned)g_n; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned int i, nvmlDevice_t* d) {
    if (!g_inited) return NVML_ERROR_UNINITIALIZED;
    if ((int)i >= g_n) return NVML_ERROR_INVALID_ARGUMENT;
    *d = (nvmlDevice_t)&g_dev[i];
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetHandleByUUID(const char* uuid, nvmlDevice_t* d) {
    for (int i = 0; i < g_n; ++i)
        if (!strcmp(uuid, g_dev[i].uuid)) { *d = (nvmlDevice_t)&g_dev[i]; return NVML_SUCCESS; }
    return NVML_ERROR_NOT_FOUND;
}
nvmlReturn_t nvmlDeviceGetIndex(nvmlDevice_t d, unsigned int* idx) { *idx = (unsigned)((mdev_t*)d)->idx; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetUUID(nvmlDevice_t d, char* uuid, unsigned int len) {
    mdev_t* m = (mdev_t*)d;
    if (m->uuid_fail) return NVML_ERROR_GPU_IS_LOST;
    snprintf(uuid, len, "%s", m->uuid);
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetName(nvmlDevice_t d, char* name, unsigned int len) { (void)d; snprintf(name, len, "NVIDIA B200"); return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t d, nvmlMemory_t* m) {
    (void)d; m->total = 192265846784ull; m->free = m->total; m->used = 0; return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetCudaComputeCapability(nvmlDevice_t d, int* major, int* minor) { (void)d; *major = 10; *minor = 0; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetPciInfo_v3(nvmlDevice_t d, nvmlPciInfo_t* p) {
    memset(p, 0, sizeof(*p));
    snprintf(p->busId, sizeof(p->busId), "00000000:%02X:00.0", 0x10 + ((mdev_t*)d)->idx);
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetNumaNodeId(nvmlDevice_t d, unsigned int* node) { (void)d; (void)node; return NVML_ERROR_NOT_SUPPORTED; }
nvmlReturn_t nvmlDeviceGetMigMode(nvmlDevice_t d, unsigned int* cur, unsigned int* pend) { (void)d; *cur = 0; *pend = 0; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceGetSupportedEventTypes(nvmlDevice_t d, unsigned long long* t) {
    int i = ((mdev_t*)d)->idx;
    if (i < MAXD && g_query_fail[i]) return NVML_ERROR_UNKNOWN;
    *t = 0xff9fULL;      /* what a B200 on driver 580 reports (profiles/box_probe_r01.txt) */
    return NVML_SUCCESS;
}
static void busy_of(int dev, int* procs, int* util) {
    *procs = 0; *util = 0;
    const char* s = getenv("MOCK_NVML_BUSY");
    while (s && *s) {
        int d = -1, p = 0, u = 0;
        if (sscanf(s, "%d:%d:%d", &d, &p, &u) >= 2 && d == dev) { *procs = p; *util = u; }
        s = strchr(s, ',');
        if (s) ++s;
    }
}
nvmlReturn_t nvmlDeviceGetComputeRunningProcesses_v3(nvmlDevice_t d, unsigned int* count, nvmlProcessInfo_t* infos) {
    int procs, util;
    busy_of(((mdev_t*)d)->idx, &procs, &util);
    if (*count < (unsigned)procs) { *count = (unsigned)procs; return NVML_ERROR_INSUFFICIENT_SIZE; }
    for (int i = 0; i < procs; ++i) { memset(&infos[i], 0, sizeof(infos[i])); infos[i].pid = 40000u + (unsigned)i; infos[i].usedGpuMemory = 1ull << 30; }
    *count = (unsigned)procs;
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetUtilizationRates(nvmlDevice_t d, nvmlUtilization_t* u) {
    int procs, util;
    busy_of(((mdev_t*)d)->idx, &procs, &util);
    u->gpu = (unsigned)util; u->memory = (unsigned)(util / 2);
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlEventSetCreate(nvmlEventSet_t* s) { *s = (nvmlEventSet_t)&g_q; return NVML_SUCCESS; }
nvmlReturn_t nvmlEventSetFree(nvmlEventSet_t s) { (void)s; return NVML_SUCCESS; }
nvmlReturn_t nvmlDeviceRegisterEvents(nvmlDevice_t d, unsigned long long types, nvmlEventSet_t s) {
    (void)s;
    int i = ((mdev_t*)d)->idx;
    if (i < MAXD && g_no_events[i]) return NVML_ERROR_NOT_SUPPORTED;
    if (i < MAXD) g_registered[i] |= types;
    return NVML_SUCCESS;
}
static long g_event_file_pos = 0;
static void drain_event_file(void) {             /* caller holds g_mu */
    const char* path = getenv("MOCK_NVML_EVENT_FILE");
    if (!path) return;
    FILE* f = fopen(path, "r");
    if (!f) return;
    if (fseek(f, g_event_file_pos, SEEK_SET) == 0) {
        char line[128];
        while (fgets(line, sizeof(line), f)) {
            if (!strchr(line, '\n')) break;       /* a line still being written: pick it up next time */
            int kind, dev;
            unsigned long long data;
            if (sscanf(line, "%d %d %llu", &kind, &dev, &data) == 3) { g_q[g_qt % 256] = (mev_t){kind, dev, data}; g_qt++; }
            g_event_file_pos = ftell(f);
        }
    }
    fclose(f);
}
nvmlReturn_t nvmlEventSetWait_v2(nvmlEventSet_t s, nvmlEventData_t* data, unsigned int timeoutms) {
    (void)s;
    pthread_mutex_lock(&g_mu);
    g_waits++;
    drain_event_file();
    if (g_qh == g_qt) {
        pthread_mutex_unlock(&g_mu);
        if (timeoutms) {                 /* MOCK_NVML_WAIT_FULL=1: block for the whole timeout like the real driver (close-vs-wait tests) */
            unsigned ms = getenv("MOCK_NVML_WAIT_FULL") ? timeoutms : (timeoutms > 2 ? 2 : timeoutms);
            struct timespec ts = {ms / 1000, (long)(ms % 1000) * 1000000L};
            nanosleep(&ts, NULL);
        }
        return NVML_ERROR_TIMEOUT;
    }
    mev_t e = g_q[g_qh % 256];
    g_qh++;
    pthread_mutex_unlock(&g_mu);
    memset(data, 0, sizeof(*data));
    data->gpuInstanceId = 0xFFFFFFFFu;
    data->computeInstanceId = 0xFFFFFFFFu;
    switch (e.kind) {
        case 0: data->device = (nvmlDevice_t)&g_dev[e.dev]; data->eventType = nvmlEventTypeXidCriticalError; data->eventData = e.data; break;
        case 1: data->device = (nvmlDevice_t)&g_dev[e.dev]; data->eventType = nvmlEventTypeDoubleBitEccError; break;
        case 2: data->device = (nvmlDevice_t)&g_dev[e.dev]; data->eventType = nvmlEventTypeSingleBitEccError; break;
        case 3: return (nvmlReturn_t)e.data;
        case 4: data->device = (nvmlDevice_t)&g_dev[g_n]; data->eventType = nvmlEventTypeXidCriticalError; data->eventData = e.data; break;
        case 5: data->device = (nvmlDevice_t)&g_dev[g_n + 1]; data->eventType = nvmlEventTypeXidCriticalError; data->eventData = e.data; break;
        default: return NVML_ERROR_UNKNOWN;
    }
    return NVML_SUCCESS;
}

/* ---- NVLink / fabric (passive cross-checks) ---------------------------------------------------- */
static unsigned long long g_nvl_kib[MAXD][4];   /* data_tx, data_rx, raw_tx, raw_rx */
void mock_nvml_add_nvlink_traffic(int dev, unsigned long long data_kib) {
    if (dev < 0 || dev >= MAXD) return;
    g_nvl_kib[dev][0] += data_kib; g_nvl_kib[dev][1] += data_kib;
    g_nvl_kib[dev][2] += data_kib + data_kib / 8; g_nvl_kib[dev][3] += data_kib + data_kib / 8;
}
nvmlReturn_t nvmlDeviceGetNvLinkState(nvmlDevice_t d, unsigned int link, nvmlEnableState_t* st) {
    if (link >= 18) return NVML_ERROR_INVALID_ARGUMENT;
    int i = ((mdev_t*)d)->idx;
    *st = NVML_FEATURE_ENABLED;
    const char* s = getenv("MOCK_NVML_LINKS_DOWN");
    while (s && *s) {
        char* e;
        long dev = strtol(s, &e, 10);
        if (*e != ':') break;
        long l = strtol(e + 1, &e, 10);
        if (dev == i && l == (long)link) *st = NVML_FEATURE_DISABLED;
        s = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetGpuFabricInfoV(nvmlDevice_t d, nvmlGpuFabricInfoV_t* fi) {
    (void)d;
    if (fi->version != nvmlGpuFabricInfo_v2) return NVML_ERROR_ARGUMENT_VERSION_MISMATCH;
    fi->state = NVML_GPU_FABRIC_STATE_COMPLETED;
    fi->status = NVML_SUCCESS;
    fi->cliqueId = 1;
    fi->healthMask = 0;
    return NVML_SUCCESS;
}
nvmlReturn_t nvmlDeviceGetFieldValues(nvmlDevice_t d, int n, nvmlFieldValue_t* v) {
    int i = ((mdev_t*)d)->idx;
    for (int k = 0; k < n; ++k) {
        int which = (int)v[k].fieldId - NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX;
        if (which < 0 || which > 3 || i >= MAXD) { v[k].nvmlReturn = NVML_ERROR_NOT_SUPPORTED; continue; }
        v[k].nvmlReturn = NVML_SUCCESS;
        v[k].valueType = NVML_VALUE_TYPE_UNSIGNED_LONG_LONG;
        v[k].value.ullVal = g_nvl_kib[i][which];
    }
    return NVML_SUCCESS;
}

// wire_host.cpp -- host codec of the compact wire format and its worker pool (see wire_host.h).
#include "wire_host.h"
#include "kosk_common.cuh"
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace kosk {

// ---------------------------------------------------------------------------------------------------------------
// 12-bit run codec.  n is even for every run of the KOSK wire layout; the scalar loops also take an odd tail element.
static void unpack12_scalar(const uint8_t *src, uint16_t *dst, size_t n)
{
    size_t i = 0;
    for (; i + 1 < n; i += 2, src += 3) {
        const uint32_t b0 = src[0], b1 = src[1], b2 = src[2];
        dst[i] = (uint16_t)(b0 | ((b1 & 0x0F) << 8));
        dst[i + 1] = (uint16_t)((b1 >> 4) | (b2 << 4));
    }
    if (i < n) dst[i] = (uint16_t)(src[0] | ((src[1] & 0x0F) << 8));
}
static uint32_t pack12_scalar(const uint16_t *src, uint8_t *dst, size_t n)
{
    uint32_t acc = 0;
    size_t i = 0;
    for (; i + 1 < n; i += 2, dst += 3) {
        const uint32_t a0 = src[i], a1 = src[i + 1];
        acc |= a0 | a1;
        dst[0] = (uint8_t)a0; dst[1] = (uint8_t)((a0 >> 8) | (a1 << 4)); dst[2] = (uint8_t)(a1 >> 4);
    }
    if (i < n) { const uint32_t a0 = src[i]; acc |= a0; dst[0] = (uint8_t)a0; dst[1] = (uint8_t)(a0 >> 8); dst[2] = 0; }
    return acc;
}

static bool g_nt = true;        // KOSK_B200_WIRE_NT=0: ordinary stores in the unpacker (default: streaming stores; the output is not read back here)

__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))
static void unpack12_vbmi(const uint8_t *src, uint16_t *dst, size_t n)
{
    alignas(64) uint8_t idx[64]; alignas(64) uint16_t sh[32];
    for (int i = 0; i < 32; i++) { const int lo = 3 * (i >> 1) + (i & 1); idx[2 * i] = (uint8_t)lo; idx[2 * i + 1] = (uint8_t)(lo + 1); sh[i] = (i & 1) ? 4 : 0; }
    const __m512i vidx = _mm512_load_si512(idx), vsh = _mm512_load_si512(sh), m12 = _mm512_set1_epi16(0x0FFF);
    const __mmask64 m48 = 0xFFFFFFFFFFFFull;
    size_t i = 0;
    // peel pairs until the destination is 64-byte aligned, so that the body can use streaming stores
    const bool nt = g_nt && (((uintptr_t)dst & 3) == 0);
    if (nt) {
        size_t peel = ((64 - ((uintptr_t)dst & 63)) & 63) / 2;
        if (peel > n) peel = n & ~(size_t)1;
        unpack12_scalar(src, dst, peel);
        i = peel; src += peel / 2 * 3;
        for (; i + 32 <= n; i += 32, src += 48) {
            __m512i x = _mm512_permutexvar_epi8(vidx, _mm512_maskz_loadu_epi8(m48, src));
            x = _mm512_and_si512(_mm512_srlv_epi16(x, vsh), m12);
            _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), x);
        }
        _mm_sfence();
    } else {
        for (; i + 32 <= n; i += 32, src += 48) {
            __m512i x = _mm512_permutexvar_epi8(vidx, _mm512_maskz_loadu_epi8(m48, src));
            x = _mm512_and_si512(_mm512_srlv_epi16(x, vsh), m12);
            _mm512_storeu_si512(dst + i, x);
        }
    }
    unpack12_scalar(src, dst + i, n - i);
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))
static uint32_t pack12_vbmi(const uint16_t *src, uint8_t *dst, size_t n)
{
    alignas(64) uint8_t idx[64]; alignas(64) uint16_t mul[32];
    for (int j = 0; j < 64; j++) idx[j] = (uint8_t)(j < 48 ? 4 * (j / 3) + j % 3 : 0);
    for (int i = 0; i < 32; i++) mul[i] = (i & 1) ? 4096 : 1;
    const __m512i vidx = _mm512_load_si512(idx), vmul = _mm512_load_si512(mul);
    const __mmask64 m48 = 0xFFFFFFFFFFFFull;
    __m512i acc = _mm512_setzero_si512();
    size_t i = 0;
    for (; i + 32 <= n; i += 32, dst += 48) {
        const __m512i x = _mm512_loadu_si512(src + i);
        acc = _mm512_or_si512(acc, x);
        const __m512i w = _mm512_madd_epi16(x, vmul);          // a0 + 4096 a1 per 32-bit lane (24 bits)
        _mm512_mask_storeu_epi8(dst, m48, _mm512_permutexvar_epi8(vidx, w));
    }
    uint32_t a = _mm512_test_epi16_mask(acc, _mm512_set1_epi16((short)0xF000)) ? 0xF000u : 0u;
    return a | pack12_scalar(src + i, dst, n - i);
}

__attribute__((target("avx2")))
static void unpack12_avx2(const uint8_t *src, uint16_t *dst, size_t n)
{
    alignas(32) uint8_t idx[32]; alignas(32) uint16_t mul[16];
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < 8; i++) { const int lo = 3 * (i >> 1) + (i & 1); idx[16 * l + 2 * i] = (uint8_t)lo; idx[16 * l + 2 * i + 1] = (uint8_t)(lo + 1); }
    for (int i = 0; i < 16; i++) mul[i] = (i & 1) ? 1 : 16;
    const __m256i vidx = _mm256_load_si256(reinterpret_cast<const __m256i *>(idx)), vmul = _mm256_load_si256(reinterpret_cast<const __m256i *>(mul));
    size_t i = 0;
    // 16 elements = 24 bytes per step; the second 16-byte load reads bytes 12..27, so stop while 28 bytes remain
    for (; i + 16 <= n && (n - i) / 2 * 3 >= 28; i += 16, src += 24) {
        const __m128i lo = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src)), hi = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 12));
        __m256i x = _mm256_shuffle_epi8(_mm256_set_m128i(hi, lo), vidx);
        x = _mm256_srli_epi16(_mm256_mullo_epi16(x, vmul), 4);      // even lanes: (v << 4) >> 4 = low 12 bits; odd lanes: v >> 4
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i), x);
    }
    unpack12_scalar(src, dst + i, n - i);
}

static int g_simd = -1;         // 2 = avx512vbmi, 1 = avx2, 0 = scalar
static void simd_init()
{
    if (g_simd >= 0) return;
    __builtin_cpu_init();
    int s = 0;
    if (__builtin_cpu_supports("avx2")) s = 1;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512vbmi")) s = 2;
    if (const char *e = getenv("KOSK_B200_WIRE_SIMD")) { const int v = atoi(e); if (v >= 0 && v < s) s = v; }     // force a lower path (tests)
    if (const char *e = getenv("KOSK_B200_WIRE_NT")) g_nt = atoi(e) != 0;
    g_simd = s;
}
const char *wire_simd_name() { simd_init(); return g_simd == 2 ? "avx512vbmi" : g_simd == 1 ? "avx2" : "scalar"; }

static void unpack12(const uint8_t *src, uint16_t *dst, size_t n)
{
    if (g_simd == 2) unpack12_vbmi(src, dst, n);
    else if (g_simd == 1) unpack12_avx2(src, dst, n);
    else unpack12_scalar(src, dst, n);
}
static uint32_t pack12(const uint16_t *src, uint8_t *dst, size_t n)
{
    return g_simd == 2 ? pack12_vbmi(src, dst, n) : pack12_scalar(src, dst, n);
}

void wire_unpack_proof(int k, const uint8_t *wire, uint8_t *pi)
{
    simd_init();
    const WireLayout W = make_wire_layout(k);
    unpack12(wire + W.w_A, reinterpret_cast<uint16_t *>(pi + W.o_A), W.nA);
    memcpy(pi + W.o_Tcomm, wire + W.w_Tcomm, (size_t)NR * 32);
    unpack12(wire + W.w_B, reinterpret_cast<uint16_t *>(pi + W.o_B), W.nB);
    memcpy(pi + W.o_comm, wire + W.w_comm, (size_t)NR * 32);
}

int wire_pack_proof(int k, const uint8_t *pi, uint8_t *wire)
{
    simd_init();
    const WireLayout W = make_wire_layout(k);
    uint32_t acc = pack12(reinterpret_cast<const uint16_t *>(pi + W.o_A), wire + W.w_A, W.nA);
    memset(wire + W.w_A + (size_t)W.nA / 2 * 3, 0, W.w_Tcomm - (size_t)W.nA / 2 * 3);
    memcpy(wire + W.w_Tcomm, pi + W.o_Tcomm, (size_t)NR * 32);
    acc |= pack12(reinterpret_cast<const uint16_t *>(pi + W.o_B), wire + W.w_B, W.nB);
    memset(wire + W.w_B + (size_t)W.nB / 2 * 3, 0, W.w_comm - W.w_B - (size_t)W.nB / 2 * 3);
    memcpy(wire + W.w_comm, pi + W.o_comm, (size_t)NR * 32);
    memset(wire + W.w_comm + (size_t)NR * 32, 0, W.wire_bytes - W.w_comm - (size_t)NR * 32);
    return (acc & 0xF000u) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
struct WireJob { int kind, k; size_t n; const uint8_t *src; uint8_t *dst; std::atomic<int> *ctr; volatile int *flag; };
struct WireGated { void *gate; std::vector<WireJob> jobs; };
struct WirePool {
    std::vector<std::thread> th;
    std::thread gate_th;                   // the only thread that waits on gates (device events): workers never block on the device
    std::mutex m;
    std::condition_variable cv, cv_gate, cv_done;
    std::deque<WireJob> q;
    std::deque<WireGated> gq;
    size_t pending = 0;                    // jobs submitted and not finished (gated ones included)
    bool stop = false;
    wire_wait_fn wait = nullptr;
    std::atomic<uint64_t> st_gate_ns{0}, st_job_ns{0}, st_jobs{0}, st_gates{0};   // instrumentation (wire_pool_stats)
    // KOSK_B200_WIRE_TRACE=<file>: (kind, id, ns) records of the pipeline, dumped when the pool is destroyed (measurement only)
    std::mutex tm; std::vector<uint64_t> trace; const char *trace_path = nullptr;
};
static inline uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void wire_pool_trace(WirePool *p, int kind, uint64_t id)
{
    if (!p || !p->trace_path) return;
    const uint64_t t = now_ns();
    std::lock_guard<std::mutex> lk(p->tm);
    p->trace.push_back((uint64_t)kind); p->trace.push_back(id); p->trace.push_back(t);
}
static int g_skip = -1;         // KOSK_B200_WIRE_DEBUG_SKIP=1 (measurement only): jobs convert nothing, to time the link side of the pipeline alone
static void wire_run_job(const WireJob &j, int bad)
{
    if (g_skip < 0) { const char *e = getenv("KOSK_B200_WIRE_DEBUG_SKIP"); g_skip = (e && atoi(e)) ? 1 : 0; }
    if (!bad && !g_skip) {
        const WireLayout W = make_wire_layout(j.k);
        for (size_t i = 0; i < j.n; i++) {
            if (j.kind == 0) wire_unpack_proof(j.k, j.src + i * W.wire_bytes, j.dst + i * W.proof_bytes);
            else bad |= wire_pack_proof(j.k, j.src + i * W.proof_bytes, j.dst + i * W.wire_bytes);
        }
    }
    if (bad && j.flag) *j.flag = 1;
}
static void wire_finish(WirePool *p, const WireJob *jobs, size_t n)
{
    {
        std::lock_guard<std::mutex> lk(p->m);
        for (size_t i = 0; i < n; i++) if (jobs[i].ctr) jobs[i].ctr->fetch_sub(1);
        p->pending -= n;
    }
    p->cv_done.notify_all();
}

static void wire_worker(WirePool *p)
{
    for (;;) {
        WireJob j;
        {
            std::unique_lock<std::mutex> lk(p->m);
            p->cv.wait(lk, [&] { return p->stop || !p->q.empty(); });
            if (p->q.empty()) return;
            j = p->q.front(); p->q.pop_front();
        }
        const uint64_t t0 = now_ns();
        wire_run_job(j, 0);
        p->st_job_ns += now_ns() - t0; p->st_jobs++;
        wire_pool_trace(p, 2, now_ns() - t0);
        wire_finish(p, &j, 1);
    }
}

// gated groups are released to the workers in submission order once their gate has fired (slices land in stream order)
static void wire_gate_thread(WirePool *p)
{
    for (;;) {
        WireGated g;
        {
            std::unique_lock<std::mutex> lk(p->m);
            p->cv_gate.wait(lk, [&] { return p->stop || !p->gq.empty(); });
            if (p->gq.empty()) return;
            g = std::move(p->gq.front()); p->gq.pop_front();
        }
        const uint64_t t0 = now_ns();
        const int bad = (g.gate && p->wait && p->wait(g.gate) != 0) ? 1 : 0;
        p->st_gate_ns += now_ns() - t0; p->st_gates++;
        wire_pool_trace(p, 1, g.jobs.size());
        if (bad) {                         // the copy failed: nothing to convert, report through the jobs' flags
            for (const WireJob &j : g.jobs) wire_run_job(j, 1);
            wire_finish(p, g.jobs.data(), g.jobs.size());
            continue;
        }
        { std::lock_guard<std::mutex> lk(p->m); for (const WireJob &j : g.jobs) p->q.push_back(j); }
        p->cv.notify_all();
    }
}

WirePool *wire_pool_create(int threads, wire_wait_fn wait)
{
    simd_init();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    WirePool *p = new WirePool;
    p->wait = wait;
    p->trace_path = getenv("KOSK_B200_WIRE_TRACE");
    for (int i = 0; i < threads; i++) p->th.emplace_back(wire_worker, p);
    if (wait) p->gate_th = std::thread(wire_gate_thread, p);
    return p;
}
void wire_pool_destroy(WirePool *p)
{
    if (!p) return;
    wire_pool_wait_all(p);
    { std::lock_guard<std::mutex> lk(p->m); p->stop = true; }
    p->cv.notify_all(); p->cv_gate.notify_all();
    for (std::thread &t : p->th) t.join();
    if (p->gate_th.joinable()) p->gate_th.join();
    if (p->trace_path && !p->trace.empty()) {
        if (FILE *f = fopen(p->trace_path, "w")) { for (size_t i = 0; i + 2 < p->trace.size(); i += 3) fprintf(f, "%llu %llu %llu\n", (unsigned long long)p->trace[i], (unsigned long long)p->trace[i + 1], (unsigned long long)p->trace[i + 2]); fclose(f); }
    }
    delete p;
}
int wire_pool_threads(const WirePool *p) { return p ? (int)p->th.size() : 0; }
void wire_pool_stats(WirePool *p, uint64_t out[4], int reset)
{
    out[0] = p->st_gate_ns; out[1] = p->st_gates; out[2] = p->st_job_ns; out[3] = p->st_jobs;
    if (reset) { p->st_gate_ns = 0; p->st_gates = 0; p->st_job_ns = 0; p->st_jobs = 0; }
}

void wire_pool_submit(WirePool *p, void *gate, int kind, int k, size_t n, size_t per_job, const uint8_t *src, uint8_t *dst, std::atomic<int> *ctr, volatile int *flag)
{
    const WireLayout W = make_wire_layout(k);
    const size_t sin = kind == 0 ? W.wire_bytes : W.proof_bytes, sout = kind == 0 ? W.proof_bytes : W.wire_bytes;
    WireGated g; g.gate = gate;
    if (per_job < 1) per_job = 1;
    for (size_t o = 0; o < n; o += per_job) g.jobs.push_back(WireJob{kind, k, per_job < n - o ? per_job : n - o, src + o * sin, dst + o * sout, ctr, flag});
    const bool gated = gate != nullptr && p->wait != nullptr;
    {
        std::lock_guard<std::mutex> lk(p->m);
        if (ctr) ctr->fetch_add((int)g.jobs.size());
        p->pending += g.jobs.size();
        if (gated) p->gq.push_back(std::move(g));
        else for (const WireJob &j : g.jobs) p->q.push_back(j);
    }
    if (gated) p->cv_gate.notify_one(); else p->cv.notify_all();
}
void wire_pool_wait_counter(WirePool *p, std::atomic<int> *ctr)
{
    std::unique_lock<std::mutex> lk(p->m);
    p->cv_done.wait(lk, [&] { return ctr->load() == 0; });
}
void wire_pool_wait_all(WirePool *p)
{
    std::unique_lock<std::mutex> lk(p->m);
    p->cv_done.wait(lk, [&] { return p->pending == 0; });
}
int wire_pool_run(WirePool *p, int kind, int k, size_t n, const uint8_t *src, uint8_t *dst)
{
    volatile int flag = 0;
    std::atomic<int> ctr{0};
    wire_pool_submit(p, nullptr, kind, k, n, n >= 8 * p->th.size() ? 2 : 1, src, dst, &ctr, &flag);
    wire_pool_wait_counter(p, &ctr);
    return flag;
}

}  // namespace kosk

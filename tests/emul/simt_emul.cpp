// simt_emul.cpp -- fiber scheduler behind simt_emul.h (test infrastructure).
#include "simt_emul.h"

#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace hop { namespace simt {
namespace {
constexpr int W = 32;
constexpr size_t STACK = 1 << 20;
struct Warp {
    ucontext_t ctx[W], main;
    char* stacks[W];
    bool finished[W];
    int cur;
    int arrived;
    unsigned long generation;
    double xchg[W], xchg2[W];
    bool pred[W];
    void (*fn)(void*);
    void* arg;
    bool broken;
};
thread_local Warp* g = nullptr;

void yield_next() {
    Warp* w = g;
    const int me = w->cur;
    for (int step = 1; step <= W; ++step) {
        const int nxt = (me + step) % W;
        if (!w->finished[nxt]) {
            if (nxt == me) return;
            w->cur = nxt;
            swapcontext(&w->ctx[me], &w->ctx[nxt]);
            return;
        }
    }
}
void trampoline() {
    Warp* w = g;
    const int me = w->cur;
    w->fn(w->arg);
    w->finished[me] = true;
    if (w->arrived != 0) w->broken = true;   // somebody is parked at a barrier this lane will never reach
    for (int step = 1; step <= W; ++step) {
        const int nxt = (me + step) % W;
        if (!w->finished[nxt] && !w->broken) {
            w->cur = nxt;
            setcontext(&w->ctx[nxt]);
        }
    }
    setcontext(&w->main);
}
}  // namespace

int lane_id() { return g->cur; }

void sync() {
    Warp* w = g;
    const unsigned long gen = w->generation;
    if (++w->arrived == W) {
        w->arrived = 0;
        ++w->generation;
        return;
    }
    while (w->generation == gen) {
        if (w->broken) setcontext(&w->main);
        yield_next();
    }
}
double shfl(double v, int src, int width) {
    Warp* w = g;
    const int me = w->cur;
    w->xchg[me] = v;
    sync();
    const double res = w->xchg[(me & ~(width - 1)) + (src & (width - 1))];
    sync();
    return res;
}
double shfl_xor(double v, int mask, int width) {
    Warp* w = g;
    const int me = w->cur;
    w->xchg[me] = v;
    sync();
    const int base = me & ~(width - 1);
    const double res = w->xchg[base + (((me - base) ^ mask) & (width - 1))];
    sync();
    return res;
}
unsigned ballot(bool p) {
    Warp* w = g;
    w->pred[w->cur] = p;
    sync();
    unsigned m = 0;
    for (int i = 0; i < W; ++i) m |= (w->pred[i] ? 1u : 0u) << i;
    sync();
    return m;
}
bool all(bool p) { return ballot(p) == 0xffffffffu; }

void dmma(double& c0, double& c1, double a, double b) {
    Warp* w = g;
    const int me = w->cur, gg = me >> 2, t = me & 3;
    w->xchg[me] = a;    // A[g][t]
    w->xchg2[me] = b;   // B[t][g]
    sync();
    double acc[2] = {c0, c1};
    for (int s = 0; s < 2; ++s) {
        const int n = 2 * t + s;
        for (int k = 0; k < 4; ++k) acc[s] = std::fma(w->xchg[(gg << 2) | k], w->xchg2[(n << 2) | k], acc[s]);
    }
    sync();
    c0 = acc[0];
    c1 = acc[1];
}

int run_warp(void (*fn)(void*), void* arg) {
    Warp* w = (Warp*)calloc(1, sizeof(Warp));
    w->fn = fn;
    w->arg = arg;
    g = w;
    for (int i = 0; i < W; ++i) {
        w->stacks[i] = (char*)malloc(STACK);
        getcontext(&w->ctx[i]);
        w->ctx[i].uc_stack.ss_sp = w->stacks[i];
        w->ctx[i].uc_stack.ss_size = STACK;
        w->ctx[i].uc_link = &w->main;
        makecontext(&w->ctx[i], (void (*)())trampoline, 0);
    }
    w->cur = 0;
    swapcontext(&w->main, &w->ctx[0]);
    int rc = w->broken ? -1 : 0;
    for (int i = 0; i < W; ++i)
        if (!w->finished[i]) rc = -1;
    for (int i = 0; i < W; ++i) free(w->stacks[i]);
    free(w);
    g = nullptr;
    return rc;
}
}}  // namespace hop::simt

// Development aid: LD_PRELOAD sampling profiler (the container has no perf/gprof that attributes correctly):
//   gcc -O2 -shared -fPIC -o /tmp/sprof.so tools/dev/sprof.c -ldl
//   SPROF_OUT=/tmp/sp.out LD_PRELOAD=/tmp/sprof.so E264_NULL_RECON=1 oracle/oracle_decode stream.264 -q; python tools/dev/sprof_report.py /tmp/sp.out 40
// LD_PRELOAD sampling profiler: SIGPROF at ~4 kHz, records RIP; at exit writes "lib offset count" lines to $SPROF_OUT.
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#include <dlfcn.h>
#define NS (1<<22)
static unsigned long samples[NS]; static volatile unsigned long ns;
static void h(int s, siginfo_t *i, void *c) { ucontext_t *u = c; unsigned long k = __sync_fetch_and_add(&ns, 1); if (k < NS) samples[k] = u->uc_mcontext.gregs[REG_RIP]; }
static int cmp(const void *a, const void *b) { unsigned long x = *(const unsigned long*)a, y = *(const unsigned long*)b; return x < y ? -1 : x > y; }
__attribute__((destructor)) static void fini(void) {
    struct itimerval z = {{0,0},{0,0}}; setitimer(ITIMER_PROF, &z, 0);
    const char *o = getenv("SPROF_OUT"); FILE *f = fopen(o ? o : "/tmp/sprof.out", "w"); unsigned long n = ns < NS ? ns : NS;
    qsort(samples, n, sizeof samples[0], cmp);
    for (unsigned long i = 0; i < n;) { unsigned long j = i; while (j < n && samples[j] == samples[i]) j++;
        Dl_info d; if (dladdr((void*)samples[i], &d) && d.dli_fname) fprintf(f, "%s %lx %lu %s\n", d.dli_fname, samples[i] - (unsigned long)d.dli_fbase, j - i, d.dli_sname ? d.dli_sname : "?"); else fprintf(f, "? %lx %lu ?\n", samples[i], j - i); i = j; }
    fclose(f); }
__attribute__((constructor)) static void init(void) {
    struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = h; sa.sa_flags = SA_SIGINFO | SA_RESTART; sigaction(SIGPROF, &sa, 0);
    struct itimerval t = {{0, 250}, {0, 250}}; setitimer(ITIMER_PROF, &t, 0); }

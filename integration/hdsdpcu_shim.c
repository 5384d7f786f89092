/*
 * integration/hdsdpcu_shim.c -- accounting shared by the hook files (see hdsdpcu_shim.h).
 *
 * For every call the reference's IPM driver makes into a hook we record
 *   wall    CLOCK_MONOTONIC around the whole hook (what the host solver waits for),
 *   device  CUDA-event time on the library stream around the same region (hdsdpcu_timer_start / _stop),
 *   host    wall time the hook itself spent in host arithmetic (bound cone, LP slack inversion, symmetrisation).
 * The report printed when the Schur object is destroyed gives seconds per Schur build (= per IPM iteration) and
 * "GPU share" = min(device, wall - host) / wall of the Schur-assembly + Cholesky categories (north-star target >= 95 %).
 * HDSDPCU_PROFILE=0 switches the brackets (and their event synchronisations) off.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "hdsdpcu.h"
#include "hdsdpcu_shim.h"

static const char *g_names[SHIM_NCAT] = {
    "S/dS assembly + Cholesky(S)", "ratio test (Lanczos)", "S^-1 + Schur assembly", "Cholesky(M)", "solves with M",
    "primal recovery / X S X", "B1 host-matrix calls"
};
static double g_wall[SHIM_NCAT], g_dev[SHIM_NCAT], g_host[SHIM_NCAT];
static long g_calls[SHIM_NCAT];
static double g_t0, g_h0;
static int g_open = 0, g_enabled = -1, g_dirty = 0;
static double g_first = 0.0;

static double now_sec( void ) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static int enabled( void ) {
    if ( g_enabled < 0 ) {
        const char *e = getenv("HDSDPCU_PROFILE");
        g_enabled = ( e && atoi(e) == 0 ) ? 0 : 1;
        g_first = now_sec();
    }
    return g_enabled;
}

void shim_prof_begin( int cat ) {
    (void) cat;
    if ( !enabled() || g_open ) { g_open += ( g_open > 0 ); return; }   /* nested brackets fold into the outer one */
    g_open = 1;
    g_t0 = now_sec();
    hdsdpcu_timer_start();
}

void shim_prof_end( int cat ) {
    if ( !enabled() ) return;
    if ( g_open > 1 ) { g_open -= 1; return; }
    if ( g_open != 1 ) return;
    double ms = 0.0;
    if ( hdsdpcu_timer_stop(&ms) != 0 ) ms = 0.0;
    g_wall[cat] += now_sec() - g_t0;
    g_dev[cat] += 1e-3 * ms;
    g_calls[cat] += 1;
    g_open = 0;
    g_dirty = 1;
}

void shim_prof_host_begin( void ) { if ( enabled() ) g_h0 = now_sec(); }
void shim_prof_host_end( int cat ) { if ( enabled() ) g_host[cat] += now_sec() - g_h0; }

void shim_prof_report( void ) {
    if ( !enabled() || !g_dirty ) return;
    g_dirty = 0;
    double W = 0, D = 0, H = 0;
    printf("\n[hdsdpcu] hot-path accounting (wall = host clock around each hook, device = CUDA events around the same region)\n");
    printf("    %-30s %8s %11s %11s %11s %7s\n", "stage", "calls", "wall [s]", "device [s]", "host [s]", "GPU %");
    for ( int c = 0; c < SHIM_NCAT; ++c ) {
        if ( !g_calls[c] ) continue;
        double d = g_dev[c];
        if ( d > g_wall[c] - g_host[c] ) d = g_wall[c] - g_host[c];
        printf("    %-30s %8ld %11.4f %11.4f %11.4f %6.1f%%\n", g_names[c], g_calls[c], g_wall[c], g_dev[c], g_host[c],
               g_wall[c] > 0 ? 100.0 * d / g_wall[c] : 0.0);
        if ( c == SHIM_CAT_SFORM || c == SHIM_CAT_SCHUR || c == SHIM_CAT_FACTOR || c == SHIM_CAT_SOLVE ) {
            W += g_wall[c]; D += d; H += g_host[c];
        }
    }
    long builds = g_calls[SHIM_CAT_FACTOR];
    double total = now_sec() - g_first;
    printf("    Schur assembly + Cholesky (S and M) + solves: wall %.4f s, device %.4f s, host arithmetic %.4f s -> GPU share %.1f%%\n",
           W, D, H, W > 0 ? 100.0 * D / W : 0.0);
    if ( builds > 0 )
        printf("    %ld factorisations of M: %.4f s hot path per factorisation (one per IPM iteration); whole run %.2f s wall since the first hook\n",
               builds, W / (double) builds, total);
    fflush(stdout);
}

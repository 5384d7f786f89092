// hdsdp_b200/csrc/dist.cu -- multi-GPU Cholesky of the Schur matrix M over NVLink peer memory.
//
// Replaces, for m large enough to matter (config D: m = 50 000), the single dpotrf behind the reference's
// HKKTFactorize (interface/hdsdp_schur.c:328, linalg/hdsdp_linsolver.c:1096).  SURVEY section 8(e).
//
// Layout.  P ranks (one process per GPU, or several "ranks" inside one process for tests).  M is cut into block
// columns of width nb; block column j belongs to rank j % P (1-D block-cyclic).  Every rank keeps a FULL mp x mp
// buffer L (20 GB at m = 50k out of 180 GB HBM): its own block columns hold M on entry (the Schur assembly writes
// them there directly, cone.cu `owns_col`) and are updated in place; the block columns of the other ranks receive
// the finished factor panels.  On exit every rank holds the complete factor, so the triangular solves that follow
// (2..26 per factorisation) run locally without any communication.
//
// Algorithm: right-looking with one step of look-ahead.  At step k every rank applies panel k to the block columns
// it owns with ONE DMMA GEMM launch (gemm_nt.cu, block-cyclic N enumeration).  The owner of block k+1 first updates
// only that block column, then factors it on a high-priority side stream (recursive potrf + panel trsm, chol.cu)
// and pushes the panel to the peers while its main stream continues with the rest of its trailing update.
//
// Transport.  No collective library on the data path.  The next owner is the only rank whose critical path waits for a
// panel, so the panel solve is FUSED with the hand-off to it: the epilogue of the triangular-solve GEMM stores every
// finished tile both locally and, over NVLink, into the next owner's L buffer (IPC-mapped peer memory; GemmArgs::peerC),
// so the transfer overlaps the math tile by tile and no copy sits on the chain.  The remaining peers (and the small
// diagonal block / inverse leaves) are served afterwards by the copy engines in ring order.  Each delivery is published
// as a monotone sequence number in the peer's control block (st.release.sys).  Consumers wait for that number with a one-thread acquire spin on their
// own GPU (bounded by a timeout).  Ranks that live in the same process are synchronised with CUDA events instead,
// which is what lets the whole schedule run -- and be tested -- with P ranks on a single GPU.
#include "common.h"
#include <algorithm>
#include <cstring>
#include <vector>

// apply panels in pairs (K = 2 nb) to the block columns that are not next in line.  Off by default: measured on 8 B200 (m = 50k,
// nb = 256) the doubled bulk launch delays the look-ahead of the following step more than the K = 512 GEMM gains
// (0.233 s against 0.203 s per iteration)
int g_dist_delay = 0;
void hd_dist_set_delay(int on) { g_dist_delay = on; }

namespace {

constexpr int MAXP = 16;
constexpr int CTRL_BYTES = 8192;
constexpr int MAIL_OFF = 4096;      // doubles: mail[2][MAXP][8]
constexpr int MAIL_CNT = 8;
constexpr long long WAIT_TIMEOUT_NS = 30ll * 1000 * 1000 * 1000;

// control block (ints): [0,P) panel sequence from src | [P,2P) ready epoch from src | [2P,3P) info from src |
//                       [3P,4P) mail sequence from src | [4P] timeout flag | [4P+1,5P+1) complete-panel sequence from src
//   panel sequence    = the part below the diagonal block of panel k has landed (all the factorisation needs)
//   complete sequence = diagonal block and inverse leaves have landed as well (needed by the solves)
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void flag_write_kernel(int *remote_flag, int value, int *remote_info, const int *local_info) {
    if (remote_info) *remote_info = *local_info;
    __threadfence_system();
    st_release_sys(remote_flag, value);
}

__global__ void flag_wait_kernel(const int *flag, int target, int *err) {
    const long long t0 = global_ns();
    while (ld_acquire_sys(flag) < target) {
        __nanosleep(100);
        if (global_ns() - t0 > WAIT_TIMEOUT_NS) { *err = 1; return; }
    }
}

struct MailDst { double *mail[MAXP]; int *flag[MAXP]; int n; };

// val[cnt] -> this rank's slot in every peer's mailbox, then publish seq
__global__ void mail_publish_kernel(MailDst dst, const double *val, int cnt, int seq) {
    const int t = threadIdx.x;
    for (int q = 0; q < dst.n; ++q)
        if (t < cnt) dst.mail[q][t] = val[t];
    __syncwarp();
    __threadfence_system();
    if (t == 0)
        for (int q = 0; q < dst.n; ++q) st_release_sys(dst.flag[q], seq);
}

// wait for every peer's mail of sequence seq, then out[q*cnt + t] = mail_q[t] (own slot from val)
__global__ void mail_collect_kernel(const char *ctrl, int P, int self, int seq, const double *val, int cnt, double *out, int *err) {
    const int *ints = (const int *) ctrl;
    const double *mail = (const double *) (ctrl + MAIL_OFF) + (size_t) (seq & 1) * MAXP * MAIL_CNT;
    const int q = threadIdx.x;
    if (q < P && q != self) {
        const long long t0 = global_ns();
        while (ld_acquire_sys(ints + 3 * P + q) < seq) {
            __nanosleep(100);
            if (global_ns() - t0 > WAIT_TIMEOUT_NS) { *err = 1; break; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * cnt; i += blockDim.x) {
        const int r = i / cnt, t = i % cnt;
        out[i] = (r == self) ? val[t] : mail[r * MAIL_CNT + t];
    }
}
// LDL^T mode: every rank computed its static-pivoting floor from the diagonal entries it owns; all use the largest
__global__ void floor_max_kernel(const double *floors, int P, double *floorp) {
    double v = floors[0];
    for (int q = 1; q < P; ++q) v = fmax(v, floors[q]);
    *floorp = v;
}
} // namespace

struct DistRank {
    int rank = 0, dev = 0;
    DenseChol *chol = nullptr;
    bool own_chol = false, own_stream = false;
    cudaStream_t st = nullptr, side = nullptr, push = nullptr;
    cudaEvent_t ev_col = nullptr, ev_diag = nullptr, ev_panel = nullptr, ev_ready = nullptr;
    std::vector<cudaEvent_t> ev_recv; // per panel, recorded by the sender (same-process delivery): below-diagonal part landed
    std::vector<cudaEvent_t> ev_full; // ... diagonal block and inverse leaves landed
    char *ctrl = nullptr;
    int *h_ctrl = nullptr;
    double *d_floors = nullptr;            // [MAXP] pivot floors of all ranks (LDL^T mode)
    struct Peer {
        DistRank *local = nullptr;
        double *L = nullptr, *Dinv = nullptr, *sgn = nullptr;
        char *ctrl = nullptr;
        bool ipc = false;
    } peer[MAXP];
    int ready_epoch_seen[MAXP] = {0};      // push stream has waited for the peer's ready flag of this epoch
    int ready_epoch_seen_side[MAXP] = {0}; // side stream likewise (fused stores)
};

struct DistChol {
    int n = 0, mp = 0, nb = 0, nblk = 0, P = 1;
    int nlocal = 0;
    DistRank *local[MAXP] = {nullptr};
    int epoch = 0, mail_seq = 0;
    bool connected = false;
};

namespace {

inline int blk_size(const DistChol *d, int k) { return (k == d->nblk - 1) ? d->mp - k * d->nb : d->nb; }
inline DistRank *local_rank(DistChol *d, int rank) {
    for (int i = 0; i < d->nlocal; ++i)
        if (d->local[i]->rank == rank) return d->local[i];
    return nullptr;
}
inline int *ctrl_ints(char *c) { return (int *) c; }
inline int panel_seq(const DistChol *d, int k) { return d->epoch * (d->nblk + 1) + k + 1; }

inline int fused_peer(const DistChol *d, const DistRank *R) { return d->P > 1 ? (R->rank + 1) % d->P : -1; }

// copy engines: whole panel to every peer except the fused one, which only misses the diagonal block; inverse leaves to all
int push_panel(DistChol *d, DistRank *R, int k) {
    const int P = d->P, nb = d->nb, mp = d->mp;
    const int s0 = k * nb, bk = blk_size(d, k);
    HD_CUDA(cudaStreamWaitEvent(R->push, R->ev_panel, 0));
    const size_t off = (size_t) s0 * mp + s0;
    const size_t leaf_off = (size_t) (s0 / HD_LEAF) * HD_LEAF * HD_LEAF;
    const int fused = fused_peer(d, R);
    for (int i = 1; i < P; ++i) {
        const int q = (R->rank + i) % P; // ring order
        DistRank::Peer &pe = R->peer[q];
        if (R->ready_epoch_seen[q] != d->epoch) {
            // the peer's buffers may still be read by its previous solves: wait until it entered this factorisation
            if (pe.local) HD_CUDA(cudaStreamWaitEvent(R->push, pe.local->ev_ready, 0));
            else HDK(flag_wait_kernel)<<<1, 1, 0, R->push>>>(ctrl_ints(R->ctrl) + P + q, d->epoch, ctrl_ints(R->ctrl) + 4 * P);
            R->ready_epoch_seen[q] = d->epoch;
        }
        double *dstL = pe.local ? pe.local->chol->L : pe.L;
        double *dstD = pe.local ? pe.local->chol->Dinv : pe.Dinv;
        const int rows = (q == fused) ? bk : mp - s0; // the fused peer already holds everything below the diagonal block
        HD_CUDA(cudaMemcpy2DAsync(dstL + off, (size_t) mp * 8, R->chol->L + off, (size_t) mp * 8, (size_t) rows * 8, bk, cudaMemcpyDefault,
                                  R->push));
        HD_CUDA(cudaMemcpyAsync(dstD + leaf_off, R->chol->Dinv + leaf_off, sizeof(double) * (size_t) (bk / HD_LEAF) * HD_LEAF * HD_LEAF,
                                cudaMemcpyDefault, R->push));
        if (R->chol->ldl && q != fused)
            HD_CUDA(cudaMemcpyAsync((pe.local ? pe.local->chol->sgn : pe.sgn) + s0, R->chol->sgn + s0, sizeof(double) * bk, cudaMemcpyDefault, R->push));
        if (pe.local) {
            if (q != fused) HD_CUDA(cudaEventRecord(pe.local->ev_recv[k], R->push));
            HD_CUDA(cudaEventRecord(pe.local->ev_full[k], R->push));
        } else {
            if (q != fused)
                HDK(flag_write_kernel)<<<1, 1, 0, R->push>>>(ctrl_ints(pe.ctrl) + R->rank, panel_seq(d, k), ctrl_ints(pe.ctrl) + 2 * P + R->rank,
                                                           R->chol->dinfo);
            HDK(flag_write_kernel)<<<1, 1, 0, R->push>>>(ctrl_ints(pe.ctrl) + 4 * P + 1 + R->rank, panel_seq(d, k), nullptr, nullptr);
        }
    }
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

int factor_panel(DistChol *d, DistRank *R, int k) {
    // side stream: Cholesky of the diagonal block of block column k (as soon as ev_diag says it is up to date) and the
    // triangular solve of the rows below it (once ev_col says the rest of the block column is up to date)
    const int nb = d->nb, mp = d->mp;
    const int s = k * nb, b = blk_size(d, k), below = mp - s - b;
    double *L = R->chol->L;
    double *leaves = R->chol->Dinv + (size_t) (s / HD_LEAF) * HD_LEAF * HD_LEAF;
    HD_CUDA(cudaStreamWaitEvent(R->side, R->ev_diag, 0));
    const bool ldl = R->chol->ldl;
    if (ldl) hd_ldl_scope(R->chol);
    int prc = hd_potrf_rec(R->side, L + (size_t) s * mp + s, mp, b, leaves, R->chol->dinfo, s);
    hd_ldl_scope(nullptr);
    HD_CALL(prc);
    HD_CUDA(cudaStreamWaitEvent(R->side, R->ev_col, 0));
    const int fused = fused_peer(d, R);
    DistRank::Peer *pf = fused >= 0 ? &R->peer[fused] : nullptr;
    if (pf && R->ready_epoch_seen_side[fused] != d->epoch) {
        if (pf->local) HD_CUDA(cudaStreamWaitEvent(R->side, pf->local->ev_ready, 0));
        else HDK(flag_wait_kernel)<<<1, 1, 0, R->side>>>(ctrl_ints(R->ctrl) + d->P + fused, d->epoch, ctrl_ints(R->ctrl) + 4 * d->P);
        R->ready_epoch_seen_side[fused] = d->epoch;
    }
    if (below > 0) {
        // panel solve fused with the hand-off: every finished tile is also stored into the next owner's buffer
        if (pf) hd_trsm_set_peer(L, pf->local ? pf->local->chol->L : pf->L);
        if (ldl) hd_ldl_scope(R->chol);
        int rc = hd_trsm_rec(R->side, L + (size_t) s * mp + s + b, mp, below, L + (size_t) s * mp + s, mp, b, leaves);
        hd_trsm_set_peer(nullptr, nullptr);
        hd_ldl_scope(nullptr);
        HD_CALL(rc);
    }
    if (pf && ldl) // the next owner needs the signs of this panel for its column update
        HD_CUDA(cudaMemcpyAsync((pf->local ? pf->local->chol->sgn : pf->sgn) + s, R->chol->sgn + s, sizeof(double) * b, cudaMemcpyDefault, R->side));
    if (pf) {
        if (pf->local) HD_CUDA(cudaEventRecord(pf->local->ev_recv[k], R->side));
        else HDK(flag_write_kernel)<<<1, 1, 0, R->side>>>(ctrl_ints(pf->ctrl) + R->rank, panel_seq(d, k), ctrl_ints(pf->ctrl) + 2 * d->P + R->rank,
                                                        R->chol->dinfo);
    }
    HD_CUDA(cudaEventRecord(R->ev_panel, R->side));
    return push_panel(d, R, k);
}

} // namespace

int dist_create(DistChol **pd, int n, int nb, int P, int nlocal, const int *ranks, DenseChol **chols, cudaStream_t main_stream) {
    if (P < 1 || P > MAXP || nlocal < 1 || nlocal > P || nb < HD_LEAF || nb % HD_LEAF) return HD_FAILED;
    DistChol *d = new DistChol();
    d->n = n; d->mp = hd_pad(n); d->nb = nb; d->P = P; d->nlocal = nlocal;
    d->nblk = (d->mp + nb - 1) / nb;
    int lo = 0, hi = 0;
    HD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int i = 0; i < nlocal; ++i) {
        DistRank *R = new DistRank();
        d->local[i] = R;
        R->rank = ranks[i];
        HD_CUDA(cudaGetDevice(&R->dev));
        if (chols && chols[i]) R->chol = chols[i];
        else { HD_CALL(chol_create(&R->chol, n)); R->own_chol = true; }
        if (main_stream && nlocal == 1) R->st = main_stream;
        else { HD_CUDA(cudaStreamCreateWithFlags(&R->st, cudaStreamNonBlocking)); R->own_stream = true; }
        HD_CUDA(cudaStreamCreateWithPriority(&R->side, cudaStreamNonBlocking, hi));
        HD_CUDA(cudaStreamCreateWithPriority(&R->push, cudaStreamNonBlocking, hi));
        HD_CUDA(cudaEventCreateWithFlags(&R->ev_col, cudaEventDisableTiming));
        HD_CUDA(cudaEventCreateWithFlags(&R->ev_diag, cudaEventDisableTiming));
        HD_CUDA(cudaEventCreateWithFlags(&R->ev_panel, cudaEventDisableTiming));
        HD_CUDA(cudaEventCreateWithFlags(&R->ev_ready, cudaEventDisableTiming));
        R->ev_recv.resize(d->nblk); R->ev_full.resize(d->nblk);
        for (auto &e : R->ev_recv) HD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto &e : R->ev_full) HD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        HD_CUDA(cudaMalloc(&R->ctrl, CTRL_BYTES));
        HD_CUDA(cudaMemset(R->ctrl, 0, CTRL_BYTES));
        HD_CUDA(cudaMallocHost(&R->h_ctrl, CTRL_BYTES));
    }
    for (int i = 0; i < nlocal; ++i)
        for (int j = 0; j < nlocal; ++j)
            if (i != j) d->local[i]->peer[d->local[j]->rank].local = d->local[j];
    d->connected = (nlocal == P);
    *pd = d;
    return HD_OK;
}

void dist_destroy(DistChol *d) {
    if (!d) return;
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        for (int q = 0; q < d->P; ++q)
            if (R->peer[q].ipc) { cudaIpcCloseMemHandle(R->peer[q].L); cudaIpcCloseMemHandle(R->peer[q].Dinv); cudaIpcCloseMemHandle(R->peer[q].ctrl); }
        if (R->own_chol) chol_destroy(R->chol);
        if (R->own_stream) cudaStreamDestroy(R->st);
        cudaStreamDestroy(R->side); cudaStreamDestroy(R->push);
        cudaEventDestroy(R->ev_col); cudaEventDestroy(R->ev_diag); cudaEventDestroy(R->ev_panel); cudaEventDestroy(R->ev_ready);
        for (auto &e : R->ev_recv) cudaEventDestroy(e);
        for (auto &e : R->ev_full) cudaEventDestroy(e);
        cudaFree(R->ctrl); cudaFreeHost(R->h_ctrl); if (R->d_floors) cudaFree(R->d_floors);
        delete R;
    }
    delete d;
}

DenseChol *dist_local_chol(DistChol *d, int li) { return d->local[li]->chol; }
cudaStream_t dist_local_stream(DistChol *d, int li) { return d->local[li]->st; }
int dist_block(const DistChol *d) { return d->nb; }

// 3 IPC handles (L, Dinv, control block) of local rank li
int dist_export(DistChol *d, int li, void *blob) {
    DistRank *R = d->local[li];
    cudaIpcMemHandle_t h[3];
    HD_CUDA(cudaIpcGetMemHandle(&h[0], R->chol->L));
    HD_CUDA(cudaIpcGetMemHandle(&h[1], R->chol->Dinv));
    HD_CUDA(cudaIpcGetMemHandle(&h[2], R->ctrl));
    memcpy(blob, h, sizeof(h));
    return HD_OK;
}

// blobs: P consecutive export blobs in rank order (the entries of local ranks are ignored)
int dist_connect(DistChol *d, const void *blobs) {
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *) blobs;
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        for (int q = 0; q < d->P; ++q) {
            if (q == R->rank || R->peer[q].local || R->peer[q].ipc) continue;
            void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
            HD_CUDA(cudaIpcOpenMemHandle(&p0, h[3 * q + 0], cudaIpcMemLazyEnablePeerAccess));
            HD_CUDA(cudaIpcOpenMemHandle(&p1, h[3 * q + 1], cudaIpcMemLazyEnablePeerAccess));
            HD_CUDA(cudaIpcOpenMemHandle(&p2, h[3 * q + 2], cudaIpcMemLazyEnablePeerAccess));
            R->peer[q].L = (double *) p0; R->peer[q].Dinv = (double *) p1; R->peer[q].ctrl = (char *) p2;
            R->peer[q].sgn = R->peer[q].Dinv + (size_t) d->mp * HD_LEAF; // sign entries follow the inverse leaves (chol_create)
            R->peer[q].ipc = true;
        }
    }
    d->connected = true;
    return HD_OK;
}

// Factor.  On entry the block columns owned by each local rank hold M (lower part, rows >= the block's first row,
// identity padded); on exit every local rank's buffer holds the complete factor and all inverse leaves.
int dist_allgather_small(DistChol *d, const double *d_val, int cnt, double *d_out);

// One static-pivoting floor for the whole matrix: the maximum of the per-rank floors (each from the owned diagonal entries).
static int dist_common_floor(DistChol *d) {
    const int P = d->P;
    if (P <= 1) return HD_OK;
    if (d->nlocal == P) { // all ranks in this process (self-test): through the host
        double mx = 0.0;
        for (int i = 0; i < P; ++i) {
            DistRank *R = d->local[i];
            double v = 0.0;
            HD_CUDA(cudaSetDevice(R->dev));
            HD_CUDA(cudaStreamSynchronize(R->st));
            HD_CUDA(cudaMemcpy(&v, R->chol->dfloor, sizeof(double), cudaMemcpyDeviceToHost));
            mx = v > mx ? v : mx;
        }
        for (int i = 0; i < P; ++i) {
            DistRank *R = d->local[i];
            HD_CUDA(cudaSetDevice(R->dev));
            HD_CUDA(cudaMemcpy(R->chol->dfloor, &mx, sizeof(double), cudaMemcpyHostToDevice));
        }
        return HD_OK;
    }
    if (d->nlocal != 1) return HD_FAILED;
    DistRank *R = d->local[0];
    if (!R->d_floors) HD_CUDA(cudaMalloc(&R->d_floors, sizeof(double) * MAXP));
    HD_CALL(dist_allgather_small(d, R->chol->dfloor, 1, R->d_floors));
    HDK(floor_max_kernel)<<<1, 1, 0, R->st>>>(R->d_floors, P, R->chol->dfloor);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

// ldl: L J L^T (signed Cholesky with Bunch-Kaufman leaves, chol.cu) instead of Cholesky; the sign vector travels with the panels
int dist_factor(DistChol *d, int *info_out, bool ldl) {
    if (!d->connected) return HD_FAILED;
    d->epoch++;
    const int P = d->P, nb = d->nb, mp = d->mp, nblk = d->nblk;
    // ---- 0. enter the factorisation: reset info, tell the peers our buffers may be written -----------------------
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        HD_CUDA(cudaSetDevice(R->dev));
        HD_CUDA(cudaMemsetAsync(R->chol->dinfo, 0, sizeof(int), R->st));
        R->chol->ldl = ldl;
        if (ldl) HD_CALL(chol_ldl_prepare(R->st, R->chol, nb, R->rank, P));
    }
    if (ldl) HD_CALL(dist_common_floor(d)); // before the events below: the side streams order themselves after them
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        HD_CUDA(cudaSetDevice(R->dev));
        HD_CUDA(cudaEventRecord(R->ev_ready, R->st));
        HD_CUDA(cudaEventRecord(R->ev_diag, R->st));
        HD_CUDA(cudaEventRecord(R->ev_col, R->st));
        for (int q = 0; q < P; ++q) {
            if (q == R->rank || R->peer[q].local) continue;
            HDK(flag_write_kernel)<<<1, 1, 0, R->st>>>(ctrl_ints(R->peer[q].ctrl) + P + R->rank, d->epoch, nullptr, nullptr);
        }
        R->chol->factored = false;
    }
    // ---- panel 0 -----------------------------------------------------------------------------------------------------
    if (DistRank *R = local_rank(d, 0)) {
        HD_CUDA(cudaSetDevice(R->dev));
        HD_CALL(factor_panel(d, R, 0));
    }
    const bool trace = getenv("HDSDPCU_TRACE") != nullptr && d->nlocal == 1;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s_) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s_); tev.push_back(e); } };
    for (int k = 0; k < nblk; ++k) {
        const int owner = k % P;
        // (1) every rank obtains panel k
        for (int i = 0; i < d->nlocal; ++i) {
            DistRank *R = d->local[i];
            HD_CUDA(cudaSetDevice(R->dev));
            mark(R->st);
            if (R->rank == owner) {
                HD_CUDA(cudaStreamWaitEvent(R->st, R->ev_panel, 0));
            } else if (R->peer[owner].local) {
                HD_CUDA(cudaStreamWaitEvent(R->st, R->ev_recv[k], 0));
            } else {
                HDK(flag_wait_kernel)<<<1, 1, 0, R->st>>>(ctrl_ints(R->ctrl) + owner, panel_seq(d, k), ctrl_ints(R->ctrl) + 4 * P);
            }
        }
        mark(d->local[0]->st);
        if (k == nblk - 1) break;
        const int s0 = k * nb, bk = blk_size(d, k);
        const int next = (k + 1) % P;
        // (2) the owner of block k+1 updates that block column first, then factors and ships it from its side streams
        if (DistRank *R = local_rank(d, next)) {
            HD_CUDA(cudaSetDevice(R->dev));
            const int s1 = (k + 1) * nb, b1 = blk_size(d, k + 1);
            const double *Pk = R->chol->L + (size_t) s0 * mp;
            // diagonal block first: its Cholesky (latency-bound, one CTA at a time) then runs on the side stream while the
            // main stream updates the rectangle below it
            GemmArgs g{};
            g.M = b1; g.N = b1; g.K = bk;
            g.A = Pk + s1; g.lda = mp; g.B = Pk + s1; g.ldb = mp; g.C = R->chol->L + (size_t) s1 * mp + s1; g.ldc = mp;
            g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER;
            if (ldl) g.ksign = R->chol->sgn + s0;
            HD_CALL(hd_gemm_nt(R->st, g));
            HD_CUDA(cudaEventRecord(R->ev_diag, R->st));
            const int below = mp - s1 - b1;
            if (below > 0) {
                g.M = below; g.flags = 0;
                g.A = Pk + s1 + b1; g.C = R->chol->L + (size_t) s1 * mp + s1 + b1;
                HD_CALL(hd_gemm_nt(R->st, g));
            }
            HD_CUDA(cudaEventRecord(R->ev_col, R->st));
            HD_CALL(factor_panel(d, R, k + 1));
        }
        // (3) everybody: the rest of the owned block columns.
        // Delayed updates (Cholesky mode): panels are applied in PAIRS so that the bulk GEMM runs with K = 2 nb (the DMMA GEMM
        // loses ~6 % at K = 256 against K = 512: the C tile is read-modified-written once per launch).  Even step e: only block
        // column e+2 receives panel e (its owner needs it before the look-ahead of step e+1); odd step e+1: all owned block
        // columns j >= e+3 receive panels e and e+1 in one launch.  The look-ahead column of (2) is unaffected.
        const bool delayed = g_dist_delay && !ldl && nblk >= 4;
        for (int i = 0; i < d->nlocal; ++i) {
            DistRank *R = d->local[i];
            HD_CUDA(cudaSetDevice(R->dev));
            GemmArgs g{};
            g.lda = mp; g.ldb = mp; g.ldc = mp; g.alpha = -1.0; g.beta = 1.0; g.flags = HD_GEMM_LOWER;
            g.bc_nb = nb; g.bc_stride = P * nb;
            if (!delayed) {
                int j0 = k + 1 + ((R->rank - (k + 1)) % P + P) % P; // first owned block > k
                if (j0 == k + 1 && R->rank == next) j0 += P;       // already done in (2)
                if (j0 >= nblk) continue;
                const int cnt = (nblk - 1 - j0) / P + 1;
                const int ms = j0 * nb;
                const double *Pk = R->chol->L + (size_t) s0 * mp;
                g.M = mp - ms; g.N = cnt * nb; g.K = bk;
                g.A = Pk + ms; g.B = Pk + ms; g.C = R->chol->L + (size_t) ms * mp + ms;
                if (ldl) g.ksign = R->chol->sgn + s0;
                HD_CALL(hd_gemm_nt(R->st, g));
            } else if (k % 2 == 0) {
                const int j = k + 2;
                if (j >= nblk || j % P != R->rank) continue;
                const int ms = j * nb;
                const double *Pk = R->chol->L + (size_t) s0 * mp;
                g.M = mp - ms; g.N = nb; g.K = bk;                  // one owned block column (the last one may be narrower: M = N there)
                if (j == nblk - 1) g.N = mp - ms;
                g.A = Pk + ms; g.B = Pk + ms; g.C = R->chol->L + (size_t) ms * mp + ms;
                HD_CALL(hd_gemm_nt(R->st, g));
            } else {
                int j0 = k + 2 + ((R->rank - (k + 2)) % P + P) % P; // first owned block >= k+2
                if (j0 >= nblk) continue;
                const int cnt = (nblk - 1 - j0) / P + 1;
                const int ms = j0 * nb;
                const int sp = (k - 1) * nb;                        // panels k-1 and k are adjacent block columns of L
                const double *Pp = R->chol->L + (size_t) sp * mp;
                g.M = mp - ms; g.N = cnt * nb; g.K = nb + bk;
                g.A = Pp + ms; g.B = Pp + ms; g.C = R->chol->L + (size_t) ms * mp + ms;
                HD_CALL(hd_gemm_nt(R->st, g));
            }
        }
    }
    if (trace) {
        DistRank *R = d->local[0];
        cudaStreamSynchronize(R->st);
        double wait = 0, work = 0, wait_own = 0;
        for (int k = 0; k < nblk; ++k) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tev[2 * k], tev[2 * k + 1]);
            if (k + 1 < nblk) cudaEventElapsedTime(&b, tev[2 * k + 1], tev[2 * k + 2]);
            wait += a; work += b;
            if (k % P == R->rank) wait_own += a;
        }
        fprintf(stderr, "[trace] dist_factor rank %d/%d nb=%d: main stream waited %.1f ms for panels (%.1f ms for its own), worked %.1f ms\n",
                R->rank, P, nb, wait, wait_own, work);
        for (auto e : tev) cudaEventDestroy(e);
    }
    // ---- finish: transposed leaves, info -----------------------------------------------------------------------------
    int info = 0, err = 0;
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        HD_CUDA(cudaSetDevice(R->dev));
        // the solves need every diagonal block and inverse leaf: wait for the last complete-panel delivery of each source
        for (int q = 0; q < P; ++q) {
            if (q == R->rank) continue;
            int last = -1;
            for (int kk = nblk - 1; kk >= 0; --kk)
                if (kk % P == q) { last = kk; break; }
            if (last < 0) continue;
            if (R->peer[q].local) HD_CUDA(cudaStreamWaitEvent(R->st, R->ev_full[last], 0));
            else HDK(flag_wait_kernel)<<<1, 1, 0, R->st>>>(ctrl_ints(R->ctrl) + 4 * P + 1 + q, panel_seq(d, last), ctrl_ints(R->ctrl) + 4 * P);
        }
        HD_CALL(hd_chol_finish(R->st, R->chol));
        HD_CUDA(cudaMemcpyAsync(R->chol->hinfo, R->chol->dinfo, sizeof(int), cudaMemcpyDeviceToHost, R->st));
        HD_CUDA(cudaMemcpyAsync(R->h_ctrl, R->ctrl, sizeof(int) * (4 * P + 1), cudaMemcpyDeviceToHost, R->st));
    }
    for (int i = 0; i < d->nlocal; ++i) {
        DistRank *R = d->local[i];
        HD_CUDA(cudaSetDevice(R->dev));
        HD_CUDA(cudaStreamSynchronize(R->st));
        HD_CUDA(cudaStreamSynchronize(R->push)); // our panels have left before the caller may touch L again
        auto take = [&](int v) { if (v > 0 && v <= d->n && (info == 0 || v < info)) info = v; };
        take(*R->chol->hinfo);
        for (int q = 0; q < P; ++q)
            if (q != R->rank && !R->peer[q].local) take(R->h_ctrl[2 * P + q]);
        if (R->h_ctrl[4 * P]) err = 1;
    }
    for (int i = 0; i < d->nlocal; ++i) d->local[i]->chol->factored = (info == 0 && !err);
    if (d->nlocal > 0) HD_CUDA(cudaSetDevice(d->local[0]->dev));
    if (info_out) *info_out = info;
    if (err) {
        fprintf(stderr, "[hdsdpcu] dist_factor: timed out waiting for a peer\n");
        return HD_FAILED;
    }
    return HD_OK;
}

// all-gather of up to 8 doubles per rank across processes (used for HKKTRegularize's global min of diag(M)).
// d_out[q * cnt + t] = value t of rank q.  Only for ranks living in different processes.
int dist_allgather_small(DistChol *d, const double *d_val, int cnt, double *d_out) {
    if (!d->connected || d->nlocal != 1 || cnt > MAIL_CNT) return HD_FAILED;
    DistRank *R = d->local[0];
    const int P = d->P, seq = ++d->mail_seq;
    MailDst dst{};
    for (int q = 0; q < P; ++q) {
        if (q == R->rank) continue;
        char *c = R->peer[q].ctrl;
        dst.mail[dst.n] = (double *) (c + MAIL_OFF) + (size_t) (seq & 1) * MAXP * MAIL_CNT + (size_t) R->rank * MAIL_CNT;
        dst.flag[dst.n] = ctrl_ints(c) + 3 * P + R->rank;
        ++dst.n;
    }
    HDK(mail_publish_kernel)<<<1, 32, 0, R->st>>>(dst, d_val, cnt, seq);
    HDK(mail_collect_kernel)<<<1, 32, 0, R->st>>>(R->ctrl, P, R->rank, seq, d_val, cnt, d_out, ctrl_ints(R->ctrl) + 4 * P);
    HD_CUDA(cudaGetLastError());
    return HD_OK;
}

/*
 * spn_host.h — C binding of the C++ host layer (streamparanumal_b200/host/: spn::comm_t,
 * spn::platform_t, spn::ogs::ogs_t), for callers that are not C++ (bench.py, the tests).
 *
 * The C++ classes mirror the reference's ogs / platform / comm API (include/ogs.hpp:216-341,
 * include/platform.hpp:60-176, include/comm.hpp:95-575); each entry point below names the
 * reference member it stands for.  Return values follow spn_b200.h (0 = success); a C++
 * exception (the reference throws libp::exception / occa::exception) is caught at this boundary,
 * its text kept for spn_host_last_error(), and SPN_ERR_HOST (-16) returned.
 */
#ifndef SPN_HOST_H
#define SPN_HOST_H

#include "spn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SPN_ERR_HOST (-16)
const char* spn_host_last_error(void);

/* ---- communicator: comm_t (include/comm.hpp), ranks = processes of one node ---------------- */
typedef struct spn_host_comm* spn_host_comm_t;
/* rendezvous_dir NULL: /tmp/spn_rdv_<uid>_<MASTER_PORT>.  Collective over all `size` ranks. */
int spn_host_comm_create(spn_host_comm_t* comm, int rank, int size, const char* rendezvous_dir);
int spn_host_comm_create_world(spn_host_comm_t* comm);   /* RANK / WORLD_SIZE / MASTER_PORT from the environment */
int spn_host_comm_destroy(spn_host_comm_t comm);
int spn_host_comm_rank(spn_host_comm_t comm, int* rank);
int spn_host_comm_size(spn_host_comm_t comm, int* size);
int spn_host_comm_barrier(spn_host_comm_t comm);                              /* comm_t::Barrier */
int spn_host_comm_allreduce_i64(spn_host_comm_t comm, long long* buf, int n, int op /*0 sum 1 max 2 min*/);
int spn_host_comm_allreduce_f64(spn_host_comm_t comm, double* buf, int n, int op);
/* device scalar allreduce over NCCL (north_star: dot/norm/rdotr across ranks); identity at one rank */
int spn_host_comm_allreduce_device_f64(spn_host_comm_t comm, double* d_buf, int n, spn_stream_t stream);

/* ---- platform: platform_t (include/platform.hpp) ---------------------------------------------- */
typedef struct spn_host_platform* spn_host_platform_t;
int spn_host_platform_create(spn_host_platform_t* platform, spn_host_comm_t comm, int device_id /* <0: LOCAL_RANK */);
int spn_host_platform_destroy(spn_host_platform_t platform);
int spn_host_platform_set_stream(spn_host_platform_t platform, spn_stream_t stream);  /* device.setStream */
int spn_host_platform_finish(spn_host_platform_t platform);                           /* device.finish   */

/* ---- mesh labels: mesh_t::SetupBoxHex3D + ConnectNodes for the box (libs/mesh/meshSetupBoxHex3D.cpp) ---- */
/* ids[nx*ny*nz*(p+1)^3]; *Nglobal = number of distinct labels over all ranks */
int spn_host_box_hex_ids(int rank, int size, int nx, int ny, int nz, int p, long long* ids, long long* Nglobal);
/* process grid of the box (libs/core/rankDecomp.cpp:55-91,166-213): out[6] = size_x, size_y, size_z, rank_x, rank_y, rank_z */
int spn_host_rank_grid(int rank, int size, int* out);

/* ---- ogs_t (include/ogs.hpp:216-341) --------------------------------------------------------------- */
typedef struct spn_host_ogs* spn_host_ogs_t;
/* ogs_t::Setup(N, ids, comm, kind, method, unique, verbose, platform).  kind: 0 Unsigned, 1 Signed, 2 Halo (halo_t::Setup:
 * ids > 0 = the owned copy of a node, ids < 0 = halo copies; only spn_host_halo_* may be called on such a handle).
 * platform NULL: host-only plan (no CUDA call), used to test the partition/exchange logic on CPU. */
int spn_host_ogs_setup(spn_host_ogs_t* ogs, int N, long long* ids, spn_host_comm_t comm, int kind, int unique,
                       int verbose, spn_host_platform_t platform);
int spn_host_ogs_free(spn_host_ogs_t ogs);
/* info[16]: N, Ngather, NgatherGlobal, NlocalT, NlocalP, NhaloT, NhaloP, gather_defined, local nnzT, local nnzN,
 *           halo nnzT, halo nnzN, local tiles, peer_exchange (1: halo data moves by P2P stores, 0: NCCL send/recv), 0, 0 */
int spn_host_ogs_info(spn_host_ogs_t ogs, long long* info);
/* exchange mode info[8]: Nsend, Nrecv, NrowsP, Nrows, NranksSend, NranksRecv, 0, 0   (mode: 0 Sym 1 NoTrans 2 Trans) */
int spn_host_ogs_exchange_info(spn_host_ogs_t ogs, int mode, long long* info);
/* host arrays of the plan (borrowed int32 pointers).  which:
 *  0 local rowStartsT  1 local colIdsT  2 local rowStartsN  3 local colIdsN
 *  4 halo  rowStartsT  5 halo  colIdsT  6 halo  rowStartsN  7 halo  colIdsN
 *  per mode: 8 sendIds 9 sendRanks 10 sendCounts 11 sendOffsets 12 recvRanks 13 recvCounts 14 recvOffsets
 *           15 cmbStarts 16 cmbIds 17 recvRows 18 rowSendStarts 19 rowSendDst 20 slotDst (peer-memory pack: send slots per halo row) */
int spn_host_ogs_array(spn_host_ogs_t ogs, int which, int mode, const int** ptr, long long* len);
int spn_host_ogs_halo_labels(spn_host_ogs_t ogs, const long long** ptr, long long* len);  /* |label| of every halo row */
/* ogs_t::SetupGlobalToLocalMapping (libs/ogs/ogsSetup.cpp:833-849): out[N] = row of every node in a gathered vector, -1 = masked */
int spn_host_ogs_global_to_local(spn_host_ogs_t ogs, int* out);

/* device operations on raw device pointers; type = spn_type, op = spn_op, trans: 0 Sym 1 NoTrans 2 Trans.
 * Kernels go to the platform's current stream (spn_host_platform_set_stream). */
int spn_host_ogs_gather_scatter(spn_host_ogs_t ogs, void* d_v, int k, int type, int op, int trans);
int spn_host_ogs_gather_scatter_start(spn_host_ogs_t ogs, void* d_v, int k, int type, int op, int trans);
int spn_host_ogs_gather_scatter_finish(spn_host_ogs_t ogs, void* d_v, int k, int type, int op, int trans);
/* host-buffer gather+scatter (the reference's host-memory overload, include/ogs.hpp:251-256): h_v is PINNED host memory
 * (spn_host_malloc) of N*k entries.  start enqueues H2D -> GatherScatter -> D2H and returns; finish blocks until h_v holds the
 * result.  Up to two operations may be in flight, so the next upload overlaps the previous download. */
int spn_host_ogs_gather_scatter_host_start(spn_host_ogs_t ogs, void* h_v, int k, int type, int op, int trans);
int spn_host_ogs_gather_scatter_host_finish(spn_host_ogs_t ogs, void* h_v);
int spn_host_ogs_gather(spn_host_ogs_t ogs, void* d_gv, const void* d_v, int k, int type, int op, int trans);
int spn_host_ogs_scatter(spn_host_ogs_t ogs, void* d_v, const void* d_gv, int k, int type, int trans);

/* halo_t::{Exchange,Combine}{,Start,Finish}(deviceMemory<T>, k) (include/ogs.hpp:347-397, libs/ogs/ogsHalo.cpp:39-400) on a handle
 * set up with kind 2: Exchange copies the owner's value to every halo copy, Combine adds every copy into the owner. */
int spn_host_halo_exchange(spn_host_ogs_t halo, void* d_v, int k, int type);
int spn_host_halo_exchange_start(spn_host_ogs_t halo, void* d_v, int k, int type);
int spn_host_halo_exchange_finish(spn_host_ogs_t halo, void* d_v, int k, int type);
int spn_host_halo_combine(spn_host_ogs_t halo, void* d_v, int k, int type);
int spn_host_halo_combine_start(spn_host_ogs_t halo, void* d_v, int k, int type);
int spn_host_halo_combine_finish(spn_host_ogs_t halo, void* d_v, int k, int type);

/* key 1: order of a label's copies for the positive pick of ogs_t::Setup (0 stable, 1 the serial reference build's std::sort order);
 * key 0 (builds with -DSPN_TIMING_PROBES only): bit mask of ogs_t::GatherScatter phases to skip.  Not part of the drop-in surface */
int spn_host_debug_set(int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* SPN_HOST_H */

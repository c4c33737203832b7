/*
 * spn_b200.h — C ABI of the B200-native (sm_100a) device backend for
 * streamParanumal's hot path (BS0–BS8).
 *
 * This header is the drop-in boundary.  Every entry point replaces one OCCA
 * kernel / device-runtime call that the reference makes on this path; the
 * reference interface each one stands in for is cited as
 * <reference file>:<line> (paths relative to the streamParanumal tree).
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no C++/torch types.
 *  - all device pointers are BORROWED (obtained from spn_malloc, or any other
 *    CUDA allocation in the same process, e.g. deviceMemory<T>::ptr() in the
 *    reference, include/memory.hpp:308-313).  The ABI never allocates results.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Kernels are enqueued on it and the call returns immediately, exactly like
 *    occa::kernel::operator() on the CUDA backend
 *    (occa/src/occa/internal/modes/cuda/kernel.cpp:74-101).
 *  - return value: 0 on success; >0 is a cudaError_t; <0 is an spn error code
 *    (see spn_status).  spn_error_string() renders either.  The C++ host layer
 *    converts non-zero into libp::exception semantics (include/utils.hpp:57-75).
 *  - zero-sized work (n==0, NrowBlocks==0) is a successful no-op, and NULL
 *    pointers are accepted in that case (reference: skipped launches at
 *    libs/ogs/ogsOperator.cpp:186,195,304,313,451,461 and null handles at P=1).
 *  - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef SPN_B200_H
#define SPN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* spn_stream_t; /* cudaStream_t */
typedef void* spn_event_t;  /* cudaEvent_t  */

typedef enum {
  SPN_SUCCESS = 0,
  SPN_ERR_INVALID_ARGUMENT = -1,
  SPN_ERR_UNSUPPORTED = -2, /* e.g. not an sm_100 device */
  SPN_ERR_ROW_TOO_LONG = -3, /* a row block exceeds spn_ogs_nodes_per_block() */
  SPN_ERR_NCCL = -4,
  SPN_ERR_NO_DEVICE = -5,
  SPN_ERR_OUT_OF_MEMORY = -6
} spn_status;

/* element type / reduction op: numeric values equal the reference enums
 * ogs::Type {Float,Double,Int32,Int64} and ogs::Op {Add,Mul,Max,Min}
 * (include/ogs.hpp:178-191). */
typedef enum { SPN_FLOAT = 0, SPN_DOUBLE = 1, SPN_INT32 = 2, SPN_INT64 = 3 } spn_type;
typedef enum { SPN_ADD = 0, SPN_MUL = 1, SPN_MAX = 2, SPN_MIN = 3 } spn_op;

const char* spn_error_string(int status);
int spn_version(void);

/* ------------------------------------------------------------------------
 * Device runtime — replaces the slice of occa::device / occa::memory /
 * occa::stream that platform_t, deviceMemory<T> and ogs touch
 * (include/platform.hpp:95-154, include/memory.hpp:284-454,
 *  occa/src/occa/internal/modes/cuda/{device,buffer,memory,stream}.cpp).
 * ---------------------------------------------------------------------- */
int spn_device_count(int* count);                 /* occa::getDeviceCount("CUDA"), include/platform.hpp:165 */
int spn_set_device(int device_id);                /* device.setup("{mode:'CUDA',device_id:..}"), libs/core/platformDeviceConfig.cpp:159 */
int spn_get_device(int* device_id);
int spn_device_sm_count(int* sms);
int spn_device_sync(void);
int spn_stream_create(spn_stream_t* stream);      /* device.createStream(), libs/ogs/ogsSetup.cpp:103-104 */
/* high != 0: highest priority the device offers — pending blocks of this stream are scheduled before those of
 * normal streams (used for the halo kernels, which must slip in between the blocks of the interior operator) */
int spn_stream_create_priority(spn_stream_t* stream, int high);
int spn_stream_destroy(spn_stream_t stream);
int spn_stream_sync(spn_stream_t stream);         /* device.finish() = cuStreamSynchronize(current), occa/src/core/device.cpp:250-254 */
int spn_malloc(void** dptr, size_t bytes);        /* cuMemAlloc, uninitialised: occa/.../cuda/buffer.cpp:65-67 */
int spn_free(void* dptr);
int spn_host_malloc(void** hptr, size_t bytes);   /* pinned host: platform.hostMalloc, include/platform.hpp:117-140 */
int spn_host_free(void* hptr);
int spn_memset(void* dptr, int value, size_t bytes, spn_stream_t stream);
/* copyFrom/copyTo; async!=0 == props["async"]=true (enqueue on stream, no host wait),
 * occa/.../cuda/memory.cpp:59-139.  bytes==0 is a no-op (include/memory.hpp:338,372,400,434). */
int spn_memcpy_h2d(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async);
int spn_memcpy_d2h(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async);
int spn_memcpy_d2d(void* dst, const void* src, size_t bytes, spn_stream_t stream, int async);
/* events: device-side timing and stream ordering without host syncs (replaces the
 * device.finish() pairs at libs/ogs/ogs.cpp:74,104,117). */
int spn_event_create(spn_event_t* ev);
int spn_event_destroy(spn_event_t ev);
int spn_event_record(spn_event_t ev, spn_stream_t stream);
int spn_event_sync(spn_event_t ev);
int spn_stream_wait_event(spn_stream_t stream, spn_event_t ev);
int spn_event_elapsed_ms(spn_event_t start, spn_event_t stop, float* ms);
/* CUDA graphs: a sequence of launches and event fork/joins enqueued between begin and end on `stream` (and on the streams
 * that wait on its events) is recorded instead of run, and replayed with one call — the launch-bound small-size tail
 * (BASELINE.json configs[4]; the reference pays 2-3 device.finish() per ogs call: libs/ogs/ogs.cpp:74,104,117). */
typedef void* spn_graph_t;  /* cudaGraphExec_t */
int spn_graph_begin_capture(spn_stream_t stream);
int spn_graph_end_capture(spn_stream_t stream, spn_graph_t* graph);
int spn_graph_launch(spn_graph_t graph, spn_stream_t stream);
int spn_graph_destroy(spn_graph_t graph);

/* ------------------------------------------------------------------------
 * BS0–BS5 streaming kernels (FP64 = dfloat, include/types.h:33-69).
 * n is the reference's `dlong N` widened to 64 bits.
 *
 * Reductions need a device workspace of spn_reduce_workspace_bytes() bytes that
 * must not be shared by launches that may run concurrently.  The caller
 * zero-fills it once (spn_memset) before first use; launches leave it re-armed.
 * Pass NULL to use a per-device internal workspace (single stream use only).
 * ---------------------------------------------------------------------- */
size_t spn_reduce_workspace_bytes(void);

/* bs0(N,a,b): empty 1x1 launch.            BS/BS0/okl/bs0.okl:27-39, call site BS/BS0/src/bs0Run.cpp:53,72 */
int spn_bs0(int N, const double* a, double* b, spn_stream_t stream);
/* bs1(N,a,b): b[i]=a[i].                    BS/BS1/okl/bs1.okl:27-34, BS/BS1/src/bs1Run.cpp:54,76 */
int spn_copy_f64(int64_t n, const double* a, double* b, spn_stream_t stream);
/* bs2(N,alpha,x,beta,y): y=alpha*x+beta*y, unfused mul/mul/add as the Serial
 * reference computes it (-O3, no FMA).      BS/BS2/okl/bs2.okl:27-36, BS/BS2/src/bs2Run.cpp:57,79 */
int spn_axpby_f64(int64_t n, double alpha, const double* x, double beta, double* y, spn_stream_t stream);
/* bs3_1+bs3_2 fused: out[0]=sum a[i]^2.     BS/BS3/okl/bs3.okl:27-102, BS/BS3/src/bs3Run.cpp:58-59,83-84 */
int spn_norm2_f64(int64_t n, const double* a, double* out, void* workspace, spn_stream_t stream);
/* bs4_1+bs4_2 fused: out[0]=sum a[i]*b[i].  BS/BS4/okl/bs4.okl:27-103, BS/BS4/src/bs4Run.cpp:58-59,83-84 */
int spn_dot_f64(int64_t n, const double* a, const double* b, double* out, void* workspace, spn_stream_t stream);
/* bs5_1+bs5_2 fused: r+=alpha*Ap (unfused mul+add); rdotr[0]=sum r_new^2.
 *                                            BS/BS5/okl/bs5.okl:27-106, BS/BS5/src/bs5Run.cpp:62-63,87-88 */
int spn_cg_update_f64(int64_t n, const double* Ap, double alpha, double* r, double* rdotr, void* workspace, spn_stream_t stream);

/* Two-launch entry points with the reference kernels' exact argument lists, for a
 * kernel_t-level binding (kernel1/kernel2 of bs3/bs4/bs5).  *_1 ignores the
 * Nblocks launch hint for its own grid, leaves the finished sum in tmp[0] and
 * +0.0 in tmp[1..Nblocks); *_2 sums tmp[0..Nblocks) into out[0] as the reference
 * second pass does. */
int spn_bs3_1(int Nblocks, int N, const double* a, double* tmp, spn_stream_t stream);
int spn_bs3_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream);
int spn_bs4_1(int Nblocks, int N, const double* a, const double* b, double* tmp, spn_stream_t stream);
int spn_bs4_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream);
int spn_bs5_1(int Nblocks, int N, const double* Ap, double alpha, double* r, double* tmp, spn_stream_t stream);
int spn_bs5_2(int Nblocks, const double* tmp, double* out, spn_stream_t stream);

/* ------------------------------------------------------------------------
 * ogs kernels (libs/ogs/okl/ogsKernels.okl).  Argument lists follow the
 * reference kernel_t calls; `type`/`op` replace the JIT defines T, OGS_OP_INIT,
 * OGS_OP (libs/ogs/ogsUtils.cpp:46-123).  All index arrays are int32 (dlong).
 *
 * Row blocks: blockRowStarts[0..NrowBlocks] partitions the rows so that no
 * block holds more than spn_ogs_nodes_per_block() non-zeros — the value the
 * host must use where the reference uses ogs::gatherNodesPerBlock=512
 * (include/ogs/ogsOperator.hpp:37-38, libs/ogs/ogsOperator.cpp:522-594).
 * Blocks built with a smaller limit (e.g. the reference's 512) are accepted.
 * Sums fold in CSR order starting from the op identity, so results are
 * bit-identical to the reference Serial kernels.
 * ---------------------------------------------------------------------- */
int spn_ogs_nodes_per_block(void);

/* gather(NrowBlocks,k,blockRowStarts,rowStarts,colIds,v,gv)
 *   gv[k+row*K] = OP_{g in row} v[k+colIds[g]*K]
 * libs/ogs/okl/ogsKernels.okl:86-122; call site libs/ogs/ogsOperator.cpp:187-202 */
int spn_ogs_gather(int NrowBlocks, int K, const int* blockRowStarts, const int* rowStarts,
                   const int* colIds, const void* v, void* gv, int type, int op, spn_stream_t stream);
/* scatter(NrowBlocks,k,blockRowStarts,rowStarts,colIds,gv,v)
 *   v[k+colIds[g]*K] = gv[k+row*K]
 * libs/ogs/okl/ogsKernels.okl:127-164; call site libs/ogs/ogsOperator.cpp:305-320 */
int spn_ogs_scatter(int NrowBlocks, int K, const int* blockRowStarts, const int* rowStarts,
                    const int* colIds, const void* gv, void* v, int type, spn_stream_t stream);
/* gatherScatter(NrowBlocks,k,blockRowStarts,gRowStarts,gColIds,sRowStarts,sColIds,v), in place.
 * blockRowStarts must bound BOTH CSRs' block sizes.
 * libs/ogs/okl/ogsKernels.okl:32-81; call site libs/ogs/ogsOperator.cpp:452-479 */
int spn_ogs_gather_scatter(int NrowBlocks, int K, const int* blockRowStarts,
                           const int* gRowStarts, const int* gColIds,
                           const int* sRowStarts, const int* sColIds,
                           void* v, int type, int op, spn_stream_t stream);
/* extract(N,k,ids,q,gatherq): gatherq[n]=q[n%K + ids[n/K]*K]
 * libs/ogs/okl/ogsKernels.okl:167-177; call site libs/ogs/ogsPairwise.cpp:115 */
int spn_ogs_extract(int N, int K, const int* ids, const void* q, void* gatherq, int type,
                    spn_stream_t stream);

/* ------------------------------------------------------------------------
 * ogs operator plan — the B200-native device form of ogsOperator_t
 * (include/ogs/ogsOperator.hpp:44-138; ctor libs/ogs/ogsOperator.cpp:597-665, which
 * uploads rowStarts/colIds and the blockRowStarts launch partition of :522-594).
 * spn_ogs_plan_create replaces that upload step: from the HOST rowStarts it cuts the
 * rows into warp tiles (<=32 non-zeros, whole rows) and stores one 16-byte descriptor
 * per tile on the device; the kernels then read colIds + descriptors instead of
 * colIds + rowStarts + blockRowStarts.  Ncols is the operator's column count
 * (ogsOperator_t::Ncols: every colId < Ncols; vectors v have Ncols*K entries); the plan
 * also stores the inverse map column -> row, which turns Scatter into a streaming
 * kernel with coalesced stores.  h_colIds (host copy of colIds) may be NULL: it is reserved
 * for layouts that regroup rows on the host (two were measured and rejected this round,
 * profiles/r01_notes.md).  d_rowStarts / d_colIds are BORROWED device
 * arrays that must outlive the plan.  The three calls are the device Gather /
 * Scatter / GatherScatter of libs/ogs/ogsOperator.cpp:176-204, 295-322, 442-481 for
 * one CSR (gather CSR == scatter CSR, i.e. every case of an Unsigned operator and
 * Sym on a Signed one); sums fold in CSR order from the op identity (bit-identical
 * to the reference Serial kernels).  Built once at setup, like the reference operator.
 * ---------------------------------------------------------------------- */
typedef struct spn_ogs_plan* spn_ogs_plan_t;
int spn_ogs_plan_create(spn_ogs_plan_t* plan, int Nrows, int Ncols, const int* h_rowStarts, const int* h_colIds,
                        const int* d_rowStarts, const int* d_colIds);
int spn_ogs_plan_destroy(spn_ogs_plan_t plan);
int spn_ogs_plan_info(spn_ogs_plan_t plan, int* Ntiles, int* Nirregular, size_t* plan_bytes);
int spn_ogs_plan_gather(spn_ogs_plan_t plan, int K, const void* v, void* gv, int type, int op, spn_stream_t stream);
int spn_ogs_plan_scatter(spn_ogs_plan_t plan, int K, const void* gv, void* v, int type, spn_stream_t stream);
int spn_ogs_plan_gather_scatter(spn_ogs_plan_t plan, int K, void* v, int type, int op, spn_stream_t stream);
/* out-of-place form: w[col] = OP_{col' in row(col)} v[col'] for every column of the operator
 * (entries of w outside the operator are left untouched) */
int spn_ogs_plan_gather_scatter_to(spn_ogs_plan_t plan, int K, const void* v, void* w, int type, int op, spn_stream_t stream);
/* tuning/experiment knobs (key 4 / 5: iterations per CTA of the gather+scatter / gather tile kernel, 0 = persistent);
 * not part of the drop-in surface */
int spn_debug_set(int key, int value);

/* ------------------------------------------------------------------------
 * Halo side of the pairwise exchange.  The reference brackets its MPI exchange with four
 * kernels — halo gather (libs/ogs/ogs.cpp:58), extract into per-neighbour order
 * (libs/ogs/ogsPairwise.cpp:115), postmpi gather of own + received partials (:155) and halo
 * scatter (libs/ogs/ogs.cpp:124).  These two entry points are the same arithmetic in two
 * launches:
 *   gather_pack:      ownBuf[k+row*K]  = OP_{g in halo row} v[k+colIds[g]*K]            row < Nrows
 *                     sendBuf[k+s*K]   = the same fold for row sendIds[s]               s < Nsend
 *   combine_scatter:  acc = identity OP own[k+row*K] (only if row < NrowsP)
 *                               OP recvBuf[k+cmbIds[j]*K], j in [cmbStarts[row],cmbStarts[row+1]) in order
 *                     sRowStarts != NULL: out[k+sColIds[g]*K] = acc for every g of the row (scatter to v)
 *                     sRowStarts == NULL: out[k+row*K] = acc              (gathered layout, ogs.cpp:276)
 * cmbIds lists the receive slots of a row in ascending source rank, which is the reference's
 * summation order (own partial first, then remote partials by rank: ogsPairwise.cpp:214-232).
 * ---------------------------------------------------------------------- */
int spn_ogs_halo_gather_pack(int Nrows, int Nsend, int K, const int* rowStarts, const int* colIds, const int* sendIds,
                             const void* v, void* ownBuf, void* sendBuf, int type, int op, spn_stream_t stream);
int spn_ogs_halo_combine_scatter(int Nrows, int NrowsP, int K, const int* cmbStarts, const int* cmbIds,
                                 const void* ownBuf, const void* recvBuf, const int* sRowStarts, const int* sColIds,
                                 void* out, int type, int op, spn_stream_t stream);

/* ------------------------------------------------------------------------
 * Peer-memory halo exchange over NVLink (one process per GPU, buffers shared with CUDA IPC).
 * Instead of packing into a local send buffer and handing it to MPI / NCCL
 * (libs/ogs/ogsPairwise.cpp:115-152), the pack kernel stores every send slot STRAIGHT INTO THE
 * DESTINATION RANK'S RECEIVE BUFFER (P2P stores through NVSwitch) and, when the whole launch has
 * stored, raises a flag word in each destination (release at system scope).  The combine kernel of
 * the destination acquires the flags of its sources before it reads the receive buffer and, when the
 * whole launch has read, raises an ACK word in each source.  No communication kernel, no SM taken
 * from the interior operator, no host synchronisation.
 *
 * Every rank owns one block  [flag word per source rank | ack word per destination rank | receive
 * half 0 | receive half 1].  Calls are numbered by an epoch kept in DEVICE memory (spn_p2p_state_t,
 * advanced by the pack kernel itself), so a call carries no host-side value and a captured CUDA
 * graph of it can be replayed.  Data of epoch e uses half e & 1.  Flow control: before a pack kernel
 * stores into half h of destination d it waits until d has acknowledged this rank's previous send
 * that used half h (state->lastSend[d][h]) — also when the neighbour relation of a mode is one-way
 * (Trans / NoTrans of a Signed operator), where nothing else would throttle a pure sender.
 *
 * spn_p2p_plan_t (one per exchange mode) and spn_p2p_state_t (one per operator) live in device memory
 * and are written by the host once at setup (spn_memcpy_h2d); state must start zeroed.
 * ---------------------------------------------------------------------- */
#define SPN_P2P_MAX_PEERS 32
#define SPN_P2P_MAX_RANKS 64
#define SPN_IPC_HANDLE_BYTES 64
typedef struct {
  int nsend, nrecv;                                         /* destinations / sources of this mode */
  int slot0[SPN_P2P_MAX_PEERS + 1];                         /* send slots [slot0[i], slot0[i+1]) go to destination i */
  int dstRank[SPN_P2P_MAX_PEERS];                           /* its rank (index into state->lastSend) */
  long long remoteOffset[SPN_P2P_MAX_PEERS];                /* slot at which my first entry lands in destination i's half */
  void* dstHalf[2][SPN_P2P_MAX_PEERS];                      /* destination i's receive halves (peer memory) */
  unsigned long long* flagRemote[SPN_P2P_MAX_PEERS];        /* in destination i's block: raised to the epoch after my stores */
  const unsigned long long* ackLocal[SPN_P2P_MAX_PEERS];    /* in MY block: raised by destination i after it read */
  const unsigned long long* flagLocal[SPN_P2P_MAX_PEERS];   /* in MY block: raised by source i */
  unsigned long long* ackRemote[SPN_P2P_MAX_PEERS];         /* in source i's block: raised to the epoch after my reads */
  const void* recvHalf[2];                                  /* my receive halves */
} spn_p2p_plan_t;
typedef struct {
  unsigned long long epoch;                                 /* calls completed on this operator */
  unsigned long long lastSend[SPN_P2P_MAX_RANKS][2];        /* per destination rank and half: epoch of my last send into it */
  unsigned int packCounter, combineCounter;                 /* last-block detection, self-resetting */
} spn_p2p_state_t;
int spn_ipc_get_handle(void* dptr, void* handle64);            /* cudaIpcGetMemHandle of a spn_malloc block   */
int spn_ipc_open_handle(const void* handle64, void** dptr);    /* map a peer's block (enables peer access)    */
int spn_ipc_close_handle(void* dptr);
/* spn_ogs_halo_gather_pack with the send slots stored into the destinations: one fold per halo row, fanned out to the
 * row's send slots rowSendDst[rowSendStarts[row] .. rowSendStarts[row+1]) (each: destination index << 24 | slot within
 * that destination's message).  Always launched on a multi-rank operator, even with nothing to send: it advances the epoch. */
int spn_ogs_halo_gather_pack_p2p(int Nrows, int K, const int* rowStarts, const int* colIds, const int* rowSendStarts,
                                 const int* rowSendDst, const void* v, void* ownBuf, const spn_p2p_plan_t* d_plan,
                                 spn_p2p_state_t* d_state, int type, int op, spn_stream_t stream);
/* spn_ogs_extract with the output stored into the destinations (Scatter NoTrans: owners send their rows);
 * slotDst[s] = destination index << 24 | slot within that destination's message */
int spn_ogs_extract_p2p(int N, int K, const int* ids, const int* slotDst, const void* q, const spn_p2p_plan_t* d_plan,
                        spn_p2p_state_t* d_state, int type, spn_stream_t stream);
/* spn_ogs_halo_combine_scatter that first waits for the sources' flags, reads the receive half of the current epoch and
 * acknowledges the sources (recvBuf is taken from the plan) */
int spn_ogs_halo_combine_scatter_p2p(int Nrows, int NrowsP, int K, const int* cmbStarts, const int* cmbIds,
                                     const void* ownBuf, const spn_p2p_plan_t* d_plan, spn_p2p_state_t* d_state,
                                     const int* sRowStarts, const int* sColIds, void* out, int type, int op, spn_stream_t stream);

/* ------------------------------------------------------------------------
 * Device-buffer communication over NCCL — replaces comm_t::{Isend,Irecv,
 * Waitall} on device pointers (libs/ogs/ogsPairwise.cpp:135-152) and adds the
 * scalar allreduce the north-star asks for after norm/dot/rdotr.
 * Lives in libspn_b200_comm.so so the kernel library carries no NCCL dependency.
 * ---------------------------------------------------------------------- */
typedef struct spn_comm* spn_comm_t;
#define SPN_UNIQUE_ID_BYTES 128
int spn_comm_unique_id(void* id128);                       /* rank 0: ncclGetUniqueId */
int spn_comm_init(spn_comm_t* comm, const void* id128, int rank, int size);
int spn_comm_destroy(spn_comm_t comm);
int spn_comm_rank(spn_comm_t comm, int* rank);
int spn_comm_size(spn_comm_t comm, int* size);
/* one grouped pairwise exchange: for r<nrecv post recv of recvBytes[r] from
 * recvRanks[r] into recvBuf+recvOffsets[r]; for s<nsend send sendBytes[s] to sendRanks[s]
 * from sendBuf+sendOffsets[s] (offsets in bytes), all on `stream`. */
int spn_comm_exchange(spn_comm_t comm,
                      int nrecv, const int* recvRanks, const size_t* recvOffsets, const size_t* recvBytes, void* recvBuf,
                      int nsend, const int* sendRanks, const size_t* sendOffsets, const size_t* sendBytes, const void* sendBuf,
                      spn_stream_t stream);
int spn_comm_allreduce_sum_f64(spn_comm_t comm, double* buf, int count, spn_stream_t stream);
int spn_comm_barrier(spn_comm_t comm, spn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPN_B200_H */

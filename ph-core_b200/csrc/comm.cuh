// comm.cuh -- what the kernels and launchers of other translation units need from the multi-GPU layer
// (comm.cu): the peer-mapped control block of the one-process-per-GPU job, the record a rank contributes
// to a one-shot cross-rank reduction, and the stream-ordered signal / wait primitives of the P2P halo path.
//
// The reference has no distributed layer; BASELINE.json's north_star partitions reductions
// (per-GPU partials combined across ranks) and the heat grid (slabs + halos).  NCCL is the contract
// fallback for both; where every rank can map its peers' memory (CUDA IPC over NVLink / NVSwitch) the
// combine and the halo happen INSIDE the compute kernels: the last block of a sharded reduction stores its
// partial into every peer's slot and folds the N slots in rank order in the same launch; the stencil kernel
// stores the planes its neighbours need straight into their ghost planes.
#pragma once
#include "ph_common.cuh"

namespace ph {

constexpr int PH_MAX_PEERS = 16;

// One rank's contribution to a cross-rank combine (64 bytes, written word by word into the peer's memory;
// word 7 is written LAST, after a system-scope fence: it carries the call number the record belongs to).
//   w[0..1] sum (up to __int128) / extremum value bits / Prefix.s     w[2..3] sum of positives / Prefix.mx
//   w[4..5] sum of negatives / Prefix.mn                              w[6]    GLOBAL flat index (INT64_MAX = empty shard)
//   w[7]    arithmetic flags (low 32 bits) | call number << 32
struct alignas(64) ReduceSlot {
  uint64_t w[8];
};

// Pinned, device-mapped host record the finishing block writes (zero-copy): the host reads it after one
// stream synchronisation, no D2H copy.
struct ReduceResult {
  uint64_t value[2];
  int64_t index;       // arg*: global flat index of the FIRST extremum; INT64_MAX when every shard is empty
  int32_t status;      // PH_RED_*
  uint32_t flags;      // arithmetic flags of every rank, read-and-cleared
  uint32_t seq;        // written LAST (after a system fence): the call number -- the host polls this word
  uint32_t _pad;
};
enum { PH_RED_OK = 0, PH_RED_OVERFLOW = 1, PH_RED_NEED_EXACT = 2, PH_RED_EMPTY = 3, PH_RED_TIMEOUT = 4 };

struct CombineArgs {
  int32_t nranks = 0;              // <= 1: single-GPU reduction, no exchange
  int32_t rank = 0;
  uint32_t seq = 0;                // call number (identical on every rank: calls are collective and ordered)
  int32_t _pad = 0;
  int64_t elems_before = 0;        // elements owned by lower ranks: local flat index -> global flat index
  ReduceSlot* my_slots = nullptr;  // [PH_MAX_PEERS] of this call's parity, in MY control block (P2P) / gather send buffer (NCCL)
  ReduceSlot* peer_slots[PH_MAX_PEERS] = {nullptr};   // the same array in every peer's control block; [0] == nullptr: NCCL transport
  ReduceResult* host_out = nullptr;
};

// Per-rank control block, allocated with cudaMalloc (IPC-exportable) and mapped by every peer.
struct CtrlBlock {
  ReduceSlot slot[2][PH_MAX_PEERS];          // [call parity][source rank]
  alignas(128) uint32_t halo_flag[2];        // last halo event my LO ([0]) / HI ([1]) neighbour completed (written by THEM)
  alignas(128) uint32_t halo_ticket[4];      // local: [0]/[1] edge blocks of the running stencil pass that finished, per side; [2] push kernel
  alignas(128) int64_t nbr_planes[2];        // local plane count (ghosts included) of my LO / HI neighbour's slab
  // strided all-to-all (ph_alltoall_strided): [0][r] rank r is ready for exchange e (its destination may be
  // overwritten), [1][r] rank r's stores of exchange e into MY memory have landed -- each written by rank r
  alignas(128) uint32_t xchg_flag[2][PH_MAX_PEERS];
  // ordered all-reduce (ph_allreduce over peer memory): rank r's arithmetic flag word at the end of its fold
  alignas(128) uint32_t ar_flags[PH_MAX_PEERS];
  alignas(128) uint32_t ar_ticket[2];         // local: blocks of the push / fold launch that have finished
};

// What the two-steps-per-pass stencil kernel needs to deliver the halo itself (heat_tma.cu): output planes
// below lo_end / from hi_begin on are ALSO stored `delta` bytes away -- the same cell of the neighbour slab's
// ghost planes, reached through the peer mapping -- and the last such block releases the neighbour's flag.
struct HeatMirror {
  int64_t lo_end = INT64_MIN;                  // planes q < lo_end  -> also at address + delta_lo (my LO neighbour's upper ghosts)
  int64_t hi_begin = INT64_MAX;                // planes q >= hi_begin -> also at address + delta_hi (my HI neighbour's lower ghosts)
  int64_t delta_lo = 0, delta_hi = 0;          // bytes
  uint32_t* flag_lo = nullptr;                 // &ctrl[lo]->halo_flag[1]  (I am its HI neighbour)
  uint32_t* flag_hi = nullptr;                 // &ctrl[hi]->halo_flag[0]
  uint32_t* ticket = nullptr;                  // my ctrl->halo_ticket
  uint32_t event = 0, lo_blocks = 0, hi_blocks = 0;
  int32_t edges_first = 0;
};

struct PeerInfo {
  bool ready = false;              // every peer's control block is mapped
  int nranks = 1, rank = 0;
  CtrlBlock* ctrl[PH_MAX_PEERS] = {nullptr};      // ctrl[rank] is my own
  ReduceSlot* gather_send = nullptr;              // NCCL transport of the record exchange
  ReduceSlot* gather_recv = nullptr;
  ReduceResult* host_result = nullptr;            // pinned + mapped
  ReduceResult* host_result_dev = nullptr;        // its device address
  uint32_t reduce_seq = 0;
  uint32_t halo_event = 0;                        // last halo event number this rank issued
  uint32_t xchg_event = 0;                        // last strided all-to-all / ordered all-reduce this rank issued (collective: same on every rank)
  char* ar_scratch = nullptr;                     // ordered all-reduce: symmetric block [nranks staging chunks | 2 result chunks]
  size_t ar_chunk_cap = 0;                        // bytes per chunk the block was sized for
};
PeerInfo& peers();

// comm.cu
int32_t comm_combine_args(CombineArgs* out, int64_t elems_before);       // next call number, slots of its parity
int32_t comm_allgather_records(cudaStream_t s);                           // gather_send[rank] of every rank -> gather_recv[0..n)
bool comm_is_multi();                                                     // ph_comm_init was called with nranks > 1

// stream-ordered wait until *flag_dev >= value (cuStreamWaitValue32 when the driver offers it, a one-thread
// spinning kernel with a time-out otherwise)
int32_t stream_wait_geq(cudaStream_t s, uint32_t* flag_dev, uint32_t value);

}  // namespace ph

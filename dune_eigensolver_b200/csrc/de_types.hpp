// Plain argument structs shared by the host runtime (de_internal.hpp) and the kernels that consume them
// (kernels_peer.cuh, kernels_tail.cuh): the NVLink peer window layout and the fused reduction tail.
#pragma once

#include <cstddef>
#include <cstdint>

namespace de
{

  constexpr int kPeerMaxRanks = 8;
  constexpr int kPeerSlotDoubles = 4224; // >= 64 + 64 * 64
  constexpr size_t kPeerFlagBytes = 4096;
  constexpr size_t kPeerArOff = kPeerFlagBytes;
  constexpr int kPeerArChannels = 2; // all-reduce channels, each with two slot / flag sets (parity of its own epoch counter)
  constexpr size_t kPeerHaloOff = kPeerArOff + (size_t)kPeerArChannels * 2 * kPeerMaxRanks * kPeerSlotDoubles * sizeof(double);

  struct PeerArgs
  {
    int rank, nranks;
    unsigned char *base[kPeerMaxRanks]; // window of every rank (own: local pointer)
    unsigned long long epoch;           // of this operation; parity = epoch & 1
    int channel;                        // all-reduces: 0 = every all-reduce that always executes, 1 = the reduction of the second
                                        // CholQR sweep, which skips itself on the device when one sweep is enough (the epochs of
                                        // a channel advance on the host whether or not the launch executes; an executed all-reduce
                                        // of channel 0 lies between any two of channel 1, so a skipped epoch cannot alias a slot)
    const int *done;                    // converged driver loop: no-op (the same on every rank)
    int *err;                           // device error flag: a peer did not arrive
    long long timeout;                  // clocks after which a spin on a peer flag gives up
  };

  struct HaloPushArgs
  {
    int npeers;
    int peer_rank[kPeerMaxRanks];
    long long send_off[kPeerMaxRanks + 1]; // rows sent to peer p: send_rows[send_off[p] .. send_off[p+1])
    long long deposit[kPeerMaxRanks];      // first row of this rank's rows in peer p's halo block
    const int *send_rows;
    const double *X;
    int m;
    size_t halo_cap_bytes;
    int *ticket;
  };

  /** rows of an updated block that neighbours need as halo rows, when they form (at most two) contiguous ranges: the
   *  block-update kernel stores those rows into the neighbours' halo buffers while it writes them (kernels_tallskinny2.cuh),
   *  so the SpMM that follows only has to release the flags. dst[p] = first deposit row in peer p's buffer, row stride M. */
  struct PushRanges
  {
    int n;
    long long lo[2], hi[2];
    double *dst[2];
    // release != 0: this launch is the last one that can change the rows, so its last CTA (ticket) -- or CTA 0, if the
    // launch skips itself -- also releases this rank's halo flag (value epoch) in the neighbours' windows
    int release;
    unsigned long long *flag[2];
    unsigned long long epoch;
    int *ticket;
    // the update kernel starts its sweep over the row tiles at the tile that holds this row, so that BOTH boundary ranges
    // of a slab (its last and its first rows) are stored at the beginning of the launch and the peer stores drain while
    // the interior is computed; with the natural order the last plane left at the very end and the launch waited for
    // NVLink (measured on 256^3, 2 GPUs: update + 60 us, more than the separate push kernel had cost)
    long long first_row;
  };

  struct PeerList
  {
    int n;
    int rank[kPeerMaxRanks];
  };

  enum
  {
    kTailNone = 0,
    kTailChol = 1, // out = Gram matrix (m x m): Rinv = inverse Cholesky factor
    kTailConv = 2  // out = [dp (m) | ...]: convergence test of the driver loop
  };                // kTailChol | kTailConv: out = [dp (m) | Gram matrix], see reduce_tail_kernel

  struct TailArgs
  {
    int kind;
    int do_allreduce;
    PeerArgs pa;
    int *ticket;
    int m;
    // Cholesky
    double *Rinv;
    int *status;
    double *info;
    int *identity_flag;
    int *done;
    int *wellcond;       // optional int[2]: chol_inverse2_body's one-sweep decision (kernels_dense.cuh)
    int *flags_identity; // ... which also sets this flag (the second sweep's last update skips itself on it)
    const int *skip;     // optional: the whole launch is a no-op when *skip != 0 (second sweep after a one-sweep decision)
    int channel;         // all-reduce channel of this tail (PeerArgs::channel)
    // convergence
    int k;
    double shift, tol;
    double *s_prev, *hist;
    int *flags;
    // second segment of the reduction (the SpMM's Rayleigh-quotient partials), placed in front of the first
    const double *partials2;
    int nparts2, len2;
  };

} // namespace de

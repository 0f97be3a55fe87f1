// Device-resident state shared by the engine kernels and the slice kernel.
#pragma once
#include "ns_math.cuh"

namespace nsb {

// State the loop would have to return to if a speculative body has to be discarded (see DevCtl::done_iter).
struct CtlSnap {
    Key key;
    long long next_idx, num_samples, iteration;
    int cur;
};

// Inputs of one body's register update, frozen by k_iter_advance so that the update can run on its own stream while
// the next body already changes the control block.
struct EpiJob {
    long long num_samples, iteration;
    double contour;
    int old_cur;  // live buffer that held the pre-merge live set (its first m rows are the discarded shell)
    int armed;    // set by k_iter_advance, cleared by the register update that consumed the job
    unsigned long long sum_new, sum_live;  // sums of n_evals over the new rows / the merged live set (k_merge_scatter)
    unsigned bar;  // arrival counter of the register update's software grid barrier
};

// Device-resident loop control (one instance per engine).
struct DevCtl {
    Key key;               // NestedSamplerState.key
    long long next_idx;    // next_sample_idx
    long long num_samples; // num_samples
    long long iteration;
    // derived per iteration by k_iter_prologue
    Key sample_key;
    Key stream_key[3];     // sample_key of the bodies whose chain streams live in buffers 0..2 (key chain is data-independent)
    long long body;        // index of the next body to run
    double contour;
    long long disc_start;  // clamped write offset of the discarded shell
    long long ph_start;    // clamped write offset of the phantom rows
    long long sender;      // sender_node_idx of the replacements
    int active;            // 0 once the register says done: every step kernel becomes a no-op
    int cur;               // which of the two live buffers is current
    int err;               // NSB200_ERR_* bits raised by device code during the run (copied into NsRegister.error_flags)
    // The register update of body b (two log-space scans over m + N elements, 75-140 us) runs on a second stream
    // next to the slice kernel of body b + 1: the loop condition is therefore known one body late.  done_iter = the
    // iteration whose register said "done" (-1 while running); a body that started before that was known is
    // speculative: it is discarded by k_rollback (control block back to snap[done_iter & 1], its dead-store rows
    // blanked), so results are exactly those of the sequential loop.
    long long done_iter;
    long long spec_disc_start, spec_ph_start;  // where the speculative body appended (set by k_rollback)
    int spec_ran;
    CtlSnap snap[2];
    EpiJob job[2];
};

struct LiveSet {
    long long *sender;
    double *U;
    double *logL_constraint;
    double *logL;
    long long *nevals;
};

struct DeadStore {
    long long *sender;
    double *logL;
    double *U;
    long long *nevals;
    unsigned char *phantom;
    long long capacity;
};

}  // namespace nsb

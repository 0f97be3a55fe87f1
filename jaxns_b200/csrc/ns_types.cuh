// Device-resident state shared by the engine kernels and the slice kernel.
#pragma once
#include "ns_math.cuh"

namespace nsb {

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

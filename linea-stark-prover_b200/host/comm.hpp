// The communicator of a sharded prove (include/lsp_b200.h: lsp_comm).  world == 1 is the single-GPU prove.
#pragma once
#include "../csrc/stark.cuh"

struct lsp_comm {
    lsp_ctx* ctx = nullptr;
    int world = 1;
    int rank = 0;          // rank of this process (NCCL mode); unused in local mode
    bool local = false;    // all ranks hosted in this process on ctx's device
    void* nccl = nullptr;  // ncclComm_t
    // reusable upload buffers of lsp_prove_air_sharded (row-major staging, column-major trace)
    lsp::Fr* up_stage = nullptr;
    lsp::Fr* up_mat = nullptr;
    size_t up_elems = 0;
};

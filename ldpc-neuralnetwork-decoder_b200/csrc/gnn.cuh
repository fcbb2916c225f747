// gnn.cuh -- message-centred GNN decoder kernels (models/message_gnn_decoder.py).
#pragma once
#include "common.cuh"
#include "tables.cuh"

struct ldpc_gnn {
    const ldpc_code* code = nullptr;
    int layers = 0, hidden = 0, types = 0;
};

// Internal host-side interfaces of the recurrent kernels (encoder / decoder).
#pragma once
#include "common.cuh"

struct PackedMeta {
    int B, T_max, sumT;
    const int* batch_sizes;   // device [T_max]
    const int* offsets;       // device [T_max+1], exclusive prefix sum of batch_sizes
    const int* lengths;       // device [B], frames per segment, descending
};

struct EncParams {
    int cell;   // 0 LSTM, 1 GRU
    int F, H, bidir;
    const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
};
struct EncGrads {
    float* w_ih[2]; float* w_hh[2]; float* b_ih[2]; float* b_hh[2];
};

int enc_init();
size_t enc_workspace_bytes(int cell, int F, int B, int sumT, int bidir, int save);
int enc_forward(const EncParams& P, const PackedMeta& pk, const float* x, int ldx, float* last_hidden, int save,
                void* ws, size_t ws_bytes, cudaStream_t stream);
int enc_backward(const EncParams& P, const EncGrads& G, const PackedMeta& pk, const float* x, int ldx,
                 const float* d_last_hidden, void* ws, size_t ws_bytes, cudaStream_t stream);

int rnn_prepare_kernel(const void* fn, size_t smem);
bool rnn_use_tc();   // tensor-core recurrences selected (default; env ABCD_RNN_TC=0 / abcd_set_rnn_tensorcore(0) -> FFMA)

struct DecParams {
    int cell;           // 0 LSTM, 1 GRU
    int F, H, M, fin;   // output size, rnn hidden, mlp hidden, feature (+speaker embedding) size
    const float* w_f2h; const float* b_f2h;                                   // [2H|H, fin]
    const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;   // LSTMCell / GRUCell
    const float* em_w1[2]; const float* em_b1[2]; const float* em_w2[2]; const float* em_b2[2];   // mean, log-variance MLPs
    const float* off_w1; const float* off_b1; const float* off_w2; const float* off_b2;           // offset predictor
};
struct DecGrads {
    float* w_f2h; float* b_f2h;
    float* w_ih; float* w_hh; float* b_ih; float* b_hh;
    float* em_w1[2]; float* em_b1[2]; float* em_w2[2]; float* em_b2[2];
    float* off_w1; float* off_b1; float* off_w2; float* off_b2;
};

int dec_init();
size_t dec_workspace_bytes(int cell, int F, int fin, int B, int sumT);
int dec_forward(const DecParams& P, const PackedMeta& pk, const float* feat_in, const float* eps, int ld_eps,
                const float* gt, int ld_gt, const float* is_offset, float* losses /*[2] device*/,
                float* out_x, float* out_mean, float* out_logvar, float* out_offset_logit,
                void* ws, size_t ws_bytes, cudaStream_t stream);
int dec_backward(const DecParams& P, const DecGrads& G, const PackedMeta& pk, const float* feat_in,
                 const float* eps, int ld_eps, const float* gt, int ld_gt, const float* is_offset, const float* g_losses /*[2] device*/,
                 float* d_feat_in, void* ws, size_t ws_bytes, cudaStream_t stream, int parts = 3);

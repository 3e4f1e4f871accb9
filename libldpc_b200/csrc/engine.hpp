// Host engine: owns the loaded code, the device copies of its tile layouts and the workspaces,
// and launches the persistent tile kernel (kernels.cuh).  C++17 host code; the CUDA runtime is only
// touched lazily so that loader / GF(2) helpers work on machines without a GPU.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ldpc_b200.h"
#include "code.hpp"

namespace b200
{
    struct DeviceLayered;    // device-resident LayeredLayout (engine.cu; layered schedule)
    struct DeviceLayout;     // device-resident TileLayout (engine.cu; erasure decoder)
    struct DeviceSegLayout;  // device-resident SegLayout (engine.cu; tile4 kernel)

    struct FrameSource
    {
        int kind = 0;                // SRC_* of kernels.cuh
        const double *d_llr = nullptr; // SRC_LLR: device [n][nc] ...
        const float *d_llr_f32 = nullptr;  // ... or the same frames as binary32
        const int8_t *d_llr_i8 = nullptr;  // ... or as quantised int8 (LLR = value * i8_scale)
        double i8_scale = 1.0;
        const uint8_t *d_bec_in = nullptr, *d_bec_cw = nullptr; // BEC decode mode
        double x = 0;                // channel parameter (snr dB or epsilon)
        uint64_t seed = 0;
        uint32_t point = 0;
        uint64_t frame0 = 0;
    };

    struct FrameSink
    {
        double *d_llr_out = nullptr;
        uint8_t *d_hard = nullptr;
        uint32_t *d_hard_bits = nullptr; // bit-packed decisions, hard_words 32-bit words per frame
        int hard_words = 0;
        uint8_t *d_bec_out = nullptr;
        int32_t *d_iters = nullptr;
        unsigned long long *d_counters = nullptr; // [5]; null -> engine scratch
        unsigned long long *d_err_log = nullptr, *d_err_count = nullptr; // per-error diagnostics log (device), capacity in records
        unsigned long long err_cap = 0;
    };

    class Engine
    {
    public:
        Engine(const std::string &pc_file, const std::string &gen_file, int device);
        ~Engine();

        HostCode H, G;
        bool has_gen = false;
        int device = -1;
        ldpc_b200_tuning tuning{};
        ldpc_b200_stats stats{};

        // Chooses / builds the layout for (precision, algorithm) under the current tuning (host only).
        const SegLayout &layout_for(int precision, int alg, int *residency, size_t *smem_bytes);

        // Launches the tile kernel over n_frames frames on `stream` (0 = engine stream).  Asynchronous unless may_block, which
        // allows the one-off timed shape trials (they run and wait for trial kernels on `stream`).
        void launch(const decoder_param &dp, const FrameSource &src, const FrameSink &sink, uint64_t n_frames, void *stream, bool may_block = false);
        // runs the one-off shape trials for (decoder type, current precision) and a job of n_frames now (blocking)
        void prepare(const decoder_param &dp, uint64_t n_frames);

        // Blocking helpers used by the C ABI
        // llr_type: LDPC_B200_LLR_F64 / _F32 / _I8 (element type of `llr`); hard_bits: bit-packed decisions, ceil(nc/32) words per frame
        void decode_batch_host(const decoder_param &dp, const void *llr, int llr_type, double llr_scale, int64_t n, double *llr_out, uint8_t *hard,
                               uint32_t *hard_bits, int32_t *iters);
        void decode_bec_host(const decoder_param &dp, const uint8_t *in, const uint8_t *cw, int64_t n, uint8_t *out, uint8_t *hard, int32_t *iters);
        void channel_host(const std::string &channel, double x, uint64_t seed, uint32_t point, uint64_t frame0, int64_t n,
                          uint8_t *cw, double *llr, uint8_t *llr_u8);
        void sim_point(const decoder_param &dp, const std::string &channel, double x, uint64_t seed, uint32_t point,
                       uint64_t frame0, uint64_t n_frames, uint64_t counters[5], float *device_ms);
        void sim_point_log(const decoder_param &dp, const std::string &channel, double x, uint64_t seed, uint32_t point, uint64_t frame0,
                           uint64_t n_frames, uint64_t counters[5], ldpc_b200_error_record *records, int64_t capacity, int64_t *n_errors);
        void sim_point_async(const decoder_param &dp, const std::string &channel, double x, uint64_t seed, uint32_t point,
                             uint64_t frame0, uint64_t n_frames, unsigned long long *d_counters, void *stream, bool may_block = false);

        // Pipelined rounds of the sweep driver: a round is launched into one of two slots (device counters, a pinned host copy,
        // an event) and collected later, so the next round is already running while the host reads the previous one.
        void round_launch(int slot, const decoder_param &dp, const std::string &channel, double x, uint64_t seed, uint32_t point, uint64_t frame0,
                          uint64_t n_frames);
        void round_collect(int slot, uint64_t counters[5]);
        // frames of one full wave of the persistent grid for this decoder type under the current tuning (rounds are sized in waves)
        uint64_t wave_frames(const decoder_param &dp, const std::string &channel);

        // M-ASK with bit-metric decoding for the AWGN sweep (M = 2: BPSK)
        void set_modulation(int M, const int *labels, const int *bit_mapper);
        int ask_M = 2;
        std::vector<int32_t> ask_labels, ask_rev, ask_bm;
        std::vector<double> ask_X;

        // layered schedule: the layering in use (empty = built-in first-fit layering, made on first use)
        void set_layers(std::vector<std::vector<int>> layers);
        const std::vector<std::vector<int>> &layers();

        // sustained shared-memory read bandwidth of the device in GB/s (LDS.128 streaming from every SM)
        double smem_probe();
        // sustained FP64 instruction rate of the device in G thread-instructions/s (DFMA streaming from every SM)
        double fp64_probe();

        void ensure_cuda();
        void *engine_stream() const { return stream_; }

    private:
        struct Config
        {
            int precision, alg, residency, lanes, fpc, threads, ctas;
            size_t smem_bytes;
            bool tm;                                  // TMEM mirror in use
            bool wide;                                // global residency: the 1-CTA-per-SM / 128-register kernel build
            bool idx16;                               // 16-bit index entries (shared-memory residency with the TMEM mirror)
            uint32_t tm_alloc_cols, tm_cols_per_warp, tm_vn_off;
        };
        Config choose(int precision, int alg, uint64_t n_frames);
        const SegLayout &get_seg_layout(int lanes, int threads, int isz = 4);
        DeviceSegLayout &device_seg_layout(int lanes, int threads, int isz = 4);
        const TileLayout &get_layout(int fpc, int threads);
        void autotune_global(int alg, const decoder_param &dp, void *stream);
        void autotune_pair(int alg, const decoder_param &dp, void *stream);
        void maybe_autotune(int alg, const decoder_param &dp, uint64_t n_frames, void *stream);
        // outcomes of the shape trials are remembered across processes: $LDPC_B200_TUNE_CACHE (a directory; "off" disables) or
        // ~/.cache/libldpc_b200, one small text file per (code, device)
        std::string tune_cache_file();
        void tune_cache_load();
        void tune_cache_store();
        bool tune_cache_loaded_ = false;
        std::string device_name_;
        void launch_layered(const decoder_param &dp, const FrameSource &src, const FrameSink &sink, uint64_t n_frames, void *stream);
        std::vector<std::vector<int>> layers_;
        std::map<std::pair<int, int>, std::unique_ptr<LayeredLayout>> lay_layouts_;
        std::map<std::pair<int, int>, std::unique_ptr<DeviceLayered>> dev_lay_;
        void launch_bec(const decoder_param &dp, const FrameSource &src, const FrameSink &sink, uint64_t n_frames, void *stream);
        DeviceLayout &device_layout(int fpc, int threads, bool idx16);
        void ensure_state(size_t bytes, void *stream);
        void release_state(void *stream);
        bool try_seg_layout(int lanes, int threads, int isz);
        static int channel_kind(const std::string &name);

        bool cuda_ready_ = false;
        int sm_count_ = 148;
        size_t smem_optin_ = 227 * 1024, smem_per_sm_ = 228 * 1024;
        std::map<std::pair<int, int>, int> pair_tuned_; // (precision, alg) -> 1 when the two-CTAs-per-SM shape won its trial (shared-memory residency)
        int force_pair_ = -1;                           // trial: -1 = use pair_tuned_, 0/1 = force
        void *stream_ = nullptr;
        void *ev0_ = nullptr, *ev1_ = nullptr;
        std::map<std::pair<int, int>, std::unique_ptr<TileLayout>> layouts_;
        std::map<std::tuple<int, int, bool>, std::unique_ptr<DeviceLayout>> dev_layouts_;
        std::map<std::tuple<int, int, int>, std::unique_ptr<SegLayout>> seg_layouts_;
        std::map<std::tuple<int, int, int>, std::unique_ptr<DeviceSegLayout>> dev_seg_layouts_;
        std::map<std::tuple<int, int, int, int, int, size_t>, int> occupancy_;
        std::map<std::pair<int, int>, std::tuple<int, int, int>> tuned_; // (precision, alg) -> autotuned (lanes, threads, wide), global residency
        bool in_autotune_ = false;
        int force_wide_ = -1; // autotune trials: -1 = use tuned_/default, 0/1 = force
        int32_t *d_bit_pos_ = nullptr, *d_punct_ = nullptr, *d_short_ = nullptr;
        int32_t *d_g_col_ptr_ = nullptr, *d_g_row_ = nullptr; // generator matrix by column (device)
        int32_t *d_bs_row_ptr_ = nullptr, *d_bs_col_ptr_ = nullptr; // bit-sliced BEC kernel
        uint16_t *d_bs_row_slot_ = nullptr, *d_bs_col_slot_ = nullptr;
        std::unique_ptr<BecSliceLayout> bs_layout_;
        uint8_t *d_bs_tx_flag_ = nullptr;
        unsigned long long *d_counters_ = nullptr;
        double *d_ask_X_ = nullptr;
        int32_t *d_ask_label_ = nullptr, *d_ask_rev_ = nullptr, *d_ask_bm_ = nullptr;
        bool ask_uploaded_ = false;
        void ask_fill(void *ask_params); // uploads the tables on first use and fills an AskParams
        unsigned long long *d_round_[2] = {nullptr, nullptr}, *h_round_[2] = {nullptr, nullptr}; // sweep rounds in flight
        void *ev_round_[2] = {nullptr, nullptr}, *ev_round0_[2] = {nullptr, nullptr}, *ev_rdone_[2] = {nullptr, nullptr};
        unsigned char *d_state_ = nullptr;
        size_t state_bytes_ = 0;
        void *ev_state_ = nullptr; // recorded after every launch that uses the state block
        bool state_used_ = false;
        // host-buffer batch decode pipeline (decode_batch_host): copy streams, per-buffer events, cached device buffers
        void *copy_in_ = nullptr, *copy_out_ = nullptr;
        void *ev_in_[2] = {nullptr, nullptr}, *ev_k_[2] = {nullptr, nullptr}, *ev_out_[2] = {nullptr, nullptr};
        void *db_in_[2] = {nullptr, nullptr}, *db_out_[2] = {nullptr, nullptr}, *db_hard_[2] = {nullptr, nullptr}, *db_it_[2] = {nullptr, nullptr};
        void *db_bits_[2] = {nullptr, nullptr};
        size_t db_in_cap_[2] = {0, 0}, db_out_cap_[2] = {0, 0}, db_hard_cap_[2] = {0, 0}, db_it_cap_[2] = {0, 0}, db_bits_cap_[2] = {0, 0};
    };

    // reference-semantics sweep driver (sim_driver.cpp)
    int run_sweep(Engine &eng, const decoder_param &dp, const channel_param &cp, const simulation_param &sp,
                  sim_results_t *results, bool *stop_flag, int rank, int world, ldpc_b200_allreduce_fn allreduce,
                  void *user, bool quiet, bool write_file, ldpc_b200_round_fn round_fn = nullptr);
} // namespace b200

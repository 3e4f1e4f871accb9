// ldpcsim — command-line front end with the reference's interface (src/sim_cpu.cpp:7-22):
//   ldpcsim [options] codefile output-file MIN MAX STEP
//   -G/--gen-matrix FILE  -i/--num-iterations N (50)  -s/--seed N (0)  -t/--num-threads N (1, ignored)
//   --channel AWGN|BSC|BEC (AWGN)  --decoding BP|BP_MS (BP)  --max-frames N (10e9)
//   --frame-error-count N (50)  --no-early-term  -h/--help  -v/--version
// Negative MIN/MAX are positionals, as with the reference's argument parser.
// B200 extras (do not exist in the reference): --precision f64|f32, --device N, --gpus N (frames of every round sharded
// over N GPUs of this process, one host thread per GPU; the counters are summed on the host, results do not depend on N).
#include <cstdio>
#include <cstdlib>
#include <array>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "engine.hpp"

namespace
{
    const char *USAGE =
        "Usage: ldpc [options] codefile output-file snr-range \n\n"
        "Positional arguments:\n"
        "codefile            \tLDPC parity-check matrix file containing all non-zero entries.\n"
        "output-file         \tResults output file.\n"
        "snr-range           \t{MIN} {MAX} {STEP}\n\n"
        "Optional arguments:\n"
        "-h --help           \tshows help message and exits\n"
        "-v --version        \tprints version information and exits\n"
        "-G --gen-matrix     \tGenerator matrix file.\n"
        "-i --num-iterations \tNumber of iterations for decoding. (Default: 50)\n"
        "-s --seed           \tRNG seed. (Default: 0)\n"
        "-t --num-threads    \tNumber of frames to be decoded in parallel. (Default: 1; ignored, the GPU batches frames)\n"
        "--channel           \tSpecifies channel: \"AWGN\", \"BSC\", \"BEC\" (Default: AWGN)\n"
        "--decoding          \tSpecifies decoding algorithm: \"BP\", \"BP_MS\" (Default: BP)\n"
        "--max-frames        \tLimit number of decoded frames.\n"
        "--frame-error-count \tMaximum frame errors for given simulation point.\n"
        "--no-early-term     \tDisable early termination for decoding.\n"
        "--precision         \tB200 only: message arithmetic \"f64\" (bit-exact with the reference, default) or \"f32\"\n"
        "--device            \tB200 only: CUDA device index (Default: 0)\n"
        "--gpus              \tB200 only: number of GPUs to shard every round of frames over, starting at --device (Default: 1, 0 = all)\n";

    bool looks_numeric(const std::string &s)
    {
        if (s.empty()) return false;
        char *end = nullptr;
        std::strtod(s.c_str(), &end);
        return end && *end == 0;
    }

    // Multi-GPU rounds: the sweep driver hands one round [frame0, frame0 + n) of a sweep point to this callback, which
    // splits it contiguously over the engines (one per GPU, one host thread each) and sums the counters.  The frame ->
    // Philox substream mapping depends only on the global frame index, so the totals do not depend on the split.
    struct MultiGpu
    {
        std::vector<std::unique_ptr<b200::Engine>> engines;
        decoder_param dp;
        std::string channel;
        uint64_t seed = 0;
    };
    int multi_gpu_round(uint32_t point, double x, uint64_t frame0, uint64_t n, uint64_t *counters, void *user)
    {
        MultiGpu &m = *static_cast<MultiGpu *>(user);
        const size_t g = m.engines.size();
        std::vector<std::array<uint64_t, 5>> part(g);
        std::vector<std::string> err(g);
        std::vector<std::thread> th;
        for (size_t i = 0; i < g; ++i)
            th.emplace_back([&, i] {
                const uint64_t lo = frame0 + n * i / g, hi = frame0 + n * (i + 1) / g;
                part[i] = {0, 0, 0, 0, 0};
                if (hi == lo) return;
                try { m.engines[i]->sim_point(m.dp, m.channel, x, m.seed, point, lo, hi - lo, part[i].data(), nullptr); }
                catch (const std::exception &e) { err[i] = e.what(); }
            });
        for (auto &t : th) t.join();
        for (size_t i = 0; i < g; ++i)
        {
            if (!err[i].empty()) { std::cout << "Error: GPU " << i << ": " << err[i] << std::endl; return 1; }
            for (int k = 0; k < 5; ++k) counters[k] += part[i][k];
        }
        return 0;
    }

    template <typename V>
    void print_list(std::ostream &os, const std::vector<V> &v)
    { // vector printer of src/core/functions.h:62-77
        os << "[";
        for (size_t i = 0; i < v.size(); ++i) os << v[i] << (i + 1 < v.size() ? ", " : "");
        os << "]";
    }
} // namespace

int main(int argc, char **argv)
{
    std::string gen, channel = "AWGN", decoding = "BP", precision = "f64";
    unsigned iterations = 50, threads = 1;
    unsigned long seed = 0, max_frames = (unsigned long)10e9, fec = 50;
    bool early_term = true;
    int device = 0, gpus = 1;
    std::vector<std::string> pos;
    try
    {
        for (int i = 1; i < argc; ++i)
        {
            const std::string a = argv[i];
            auto value = [&]() -> std::string
            {
                if (i + 1 >= argc) throw std::runtime_error("Too few arguments for '" + a + "'.");
                return argv[++i];
            };
            if (a == "-h" || a == "--help") { std::cout << USAGE; return 0; }
            else if (a == "-v" || a == "--version") { std::cout << ldpc_b200_version() << std::endl; return 0; }
            else if (a == "-G" || a == "--gen-matrix") gen = value();
            else if (a == "-i" || a == "--num-iterations") iterations = (unsigned)std::stoul(value());
            else if (a == "-s" || a == "--seed") seed = std::stoul(value());
            else if (a == "-t" || a == "--num-threads") threads = (unsigned)std::stoul(value());
            else if (a == "--channel") channel = value();
            else if (a == "--decoding") decoding = value();
            else if (a == "--max-frames") max_frames = std::stoul(value());
            else if (a == "--frame-error-count") fec = std::stoul(value());
            else if (a == "--no-early-term") early_term = false;
            else if (a == "--precision") precision = value();
            else if (a == "--device") device = std::stoi(value());
            else if (a == "--gpus") gpus = std::stoi(value());
            else if (a.size() > 1 && a[0] == '-' && !looks_numeric(a)) throw std::runtime_error("Unknown argument: " + a);
            else pos.push_back(a);
        }
        if (pos.size() < 5) throw std::runtime_error("Too few arguments");
        if (pos.size() > 5) throw std::runtime_error("Maximum number of positional arguments exceeded");
        const double snr[3] = {std::stod(pos[2]), std::stod(pos[3]), std::stod(pos[4])};
        if (snr[0] > snr[1]) throw std::runtime_error("snr min > snr max");
        if (precision != "f64" && precision != "f32") throw std::runtime_error("--precision must be f64 or f32");

        std::unique_ptr<b200::Engine> eng;
        try
        {
            eng = std::make_unique<b200::Engine>(pos[0], gen, device);
        }
        catch (const std::exception &e)
        {
            std::cout << "Error: ldpc_code(): " << e.what() << std::endl; // src/core/ldpc.cpp:16-20
            return EXIT_FAILURE;
        }
        eng->tuning.precision = precision == "f32" ? LDPC_B200_F32 : LDPC_B200_F64;
        const auto &H = eng->H;
        const char *bar = "========================================================================================";
        std::cout << bar << std::endl;
        std::cout << "Parity-Check Matrix: " << pos[0] << std::endl;
        std::cout << "Generator Matrix: " << gen << std::endl;
        // code summary, src/core/ldpc.cpp:112-130
        std::cout << "N : " << H.nc << "\nM : " << H.mc << "\nK : " << H.kc() << "\nNNZ : " << H.nnz << "\n";
        std::cout << "puncture[" << H.puncture.size() << "] : ";
        print_list(std::cout, H.puncture);
        std::cout << "\nshorten[" << H.shorten.size() << "] : ";
        print_list(std::cout, H.shorten);
        std::cout << "\nRate : " << 1. - (double)H.mct() / (double)H.nct() << "\n";
        std::cout << "N (transmitted) : " << H.nct() << "\nM (transmitted) : " << H.mct() << "\nK (transmitted) : " << H.kct() << "\n" << std::endl;
        std::cout << bar << std::endl;

        decoder_param dp{early_term, iterations, decoding.c_str()};
        channel_param cp{seed, {snr[0], snr[1], snr[2]}, channel.c_str()};
        simulation_param sp{threads, max_frames, fec, pos[1].c_str()};
        // parameter dump, src/core/functions.cpp:19-42 and src/sim/ldpcsim.cpp:84-95
        std::cout << "== Decoder Parameters\n Type: " << dp.type << "\n Iterations: " << dp.iterations << "\n Early Termination: " << dp.earlyTerm << "\n";
        std::cout << "== Channel Parameters\n Type: " << cp.type << "\n Seed: " << cp.seed << "\n Range: Min: " << cp.xRange[0] << ", Max: " << cp.xRange[1]
                  << ", Step: " << cp.xRange[2] << "\n";
        std::cout << "== Simulation Parameters\n Threads: " << sp.threads << "\n FEC: " << sp.fec << "\n Max Frames: " << sp.maxFrames
                  << "\n Output File: " << sp.resultFile << "\n" << std::endl;
        std::cout << bar << std::endl;
        if (eng->has_gen && channel == "BEC")
            std::cout << "note: the erasure-channel sweep transmits the all-zero codeword (the generator matrix serves AWGN / BSC sweeps)" << std::endl;

        bool stop = false;
        try
        {
            if (gpus == 0 && cudaGetDeviceCount(&gpus) != cudaSuccess) gpus = 1;
            if (gpus < 1) throw std::runtime_error("--gpus must be >= 0");
            if (gpus == 1) b200::run_sweep(*eng, dp, cp, sp, nullptr, &stop, 0, 1, nullptr, nullptr, false, true);
            else
            {
                MultiGpu m;
                m.dp = dp; m.channel = channel; m.seed = seed;
                m.engines.push_back(std::move(eng));
                for (int g = 1; g < gpus; ++g)
                {
                    m.engines.push_back(std::make_unique<b200::Engine>(pos[0], gen, device + g));
                    m.engines.back()->tuning.precision = m.engines[0]->tuning.precision;
                }
                std::cout << "GPUs: " << gpus << " (devices " << device << " .. " << device + gpus - 1 << ")" << std::endl;
                b200::run_sweep(*m.engines[0], dp, cp, sp, nullptr, &stop, 0, 1, nullptr, &m, false, true, multi_gpu_round);
            }
        }
        catch (const std::exception &e)
        {
            std::cout << "Error: ldpc_sim::ldpc_sim() " << e.what() << std::endl; // src/sim/ldpcsim.cpp:77-81
            return EXIT_FAILURE;
        }
    }
    catch (const std::exception &e)
    {
        std::cout << e.what() << std::endl; // src/sim_cpu.cpp:77-82
        std::cout << USAGE;
        return EXIT_FAILURE;
    }
    return 0;
}

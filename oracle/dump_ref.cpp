// Test infrastructure (NOT product code): drives the UNMODIFIED reference classes
// (channel_* -> ldpc_decoder / ldpc_decoder_bec, /root/reference/src/sim/channel.cpp,
// /root/reference/src/decoding/decoder.cpp) and dumps their inputs/outputs so the C restatement
// (oracle/ldpc_oracle.c) and the CUDA path can be pinned against the reference itself.
// Built only in the container that has /root/reference (see oracle/Makefile); writes to oracle/_ref/.
//
// Usage:
//   dump_ref sim    H G|- AWGN|BSC|BEC BP|BP_MS iters et x seed nframes out.bin
//   dump_ref decode H BP|BP_MS iters et in.bin nframes out.bin
//
// "sim"   : per frame, in the order the reference's frame loop uses (ldpcsim.cpp:160-174):
//             [encode_and_map] -> simulate -> calculate_llrs -> decode
//           record = codeword u8[nc] | llr_in (f64[nc], or u8[nc] for BEC) | llr_out (same type)
//                    | CO u8[nc] | iters i32
// "decode": reads f64[nframes][nc] full-length LLRs (punctured/shortened positions included),
//           record = llr_out f64[nc] | CO u8[nc] | iters i32
// File header (both modes): i32 nc, i32 mc, i32 nnz, i32 nframes, i32 is_bec.
#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <forward_list>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <variant>
#include <vector>

#define private public
#define protected public
#include "sim/channel.h"
#undef private
#undef protected

using namespace ldpc;

template <typename T>
static void put(FILE *f, const T *p, size_t n) { fwrite(p, sizeof(T), n, f); }

static void put_bits(FILE *f, const vec_bits_t &v)
{
    std::vector<u8> b(v.size());
    for (size_t i = 0; i < v.size(); ++i) b[i] = v[i].value;
    put(f, b.data(), b.size());
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: see header\n"); return 2; }
    std::string mode = argv[1];
    if (mode == "sim" && argc == 12)
    {
        std::string H = argv[2], G = argv[3], ch = argv[4], dec = argv[5];
        if (G == "-") G = "";
        decoder_param dp; dp.iterations = std::stoul(argv[6]); dp.earlyTerm = std::stoi(argv[7]) != 0; dp.type = dec.c_str();
        double x = std::stod(argv[8]); u64 seed = std::stoul(argv[9]); int nframes = std::stoi(argv[10]);
        auto code = std::make_shared<ldpc_code>(H, G);
        FILE *f = fopen(argv[11], "wb");
        int32_t hdr[5] = {code->nc(), code->mc(), code->nnz(), nframes, ch == "BEC"};
        put(f, hdr, 5);
        if (ch == "AWGN")
        {
            channel_awgn c(code, dp, seed, 1.);
            c.set_channel_param(x);
            for (int t = 0; t < nframes; ++t)
            {
                if (!code->G().empty()) c.encode_and_map();
                c.simulate(); c.calculate_llrs();
                std::vector<double> in = c.mLdpcDecoder->mLLRIn;
                int32_t it = c.decode();
                put_bits(f, c.codeword()); put(f, in.data(), in.size());
                put(f, c.mLdpcDecoder->mLLROut.data(), in.size()); put_bits(f, c.estimate()); put(f, &it, 1);
            }
        }
        else if (ch == "BSC")
        {
            channel_bsc c(code, dp, seed, 0.);
            c.set_channel_param(x);
            for (int t = 0; t < nframes; ++t)
            {
                if (!code->G().empty()) c.encode_and_map();
                c.simulate(); c.calculate_llrs();
                std::vector<double> in = c.mLdpcDecoder->mLLRIn;
                int32_t it = c.decode();
                put_bits(f, c.codeword()); put(f, in.data(), in.size());
                put(f, c.mLdpcDecoder->mLLROut.data(), in.size()); put_bits(f, c.estimate()); put(f, &it, 1);
            }
        }
        else if (ch == "BEC")
        {
            channel_bec c(code, dp, seed, 0.);
            c.set_channel_param(x);
            for (int t = 0; t < nframes; ++t)
            {
                if (!code->G().empty()) c.encode_and_map();
                c.simulate(); c.calculate_llrs();
                std::vector<u8> in = c.mLdpcDecoder->mLLRIn;
                int32_t it = c.decode();
                put_bits(f, c.codeword()); put(f, in.data(), in.size());
                put(f, c.mLdpcDecoder->mLLROut.data(), in.size()); put_bits(f, c.estimate()); put(f, &it, 1);
            }
        }
        else { fprintf(stderr, "unknown channel\n"); return 2; }
        fclose(f);
        return 0;
    }
    if (mode == "decode" && argc == 9)
    {
        std::string H = argv[2], dec = argv[3];
        decoder_param dp; dp.iterations = std::stoul(argv[4]); dp.earlyTerm = std::stoi(argv[5]) != 0; dp.type = dec.c_str();
        int nframes = std::stoi(argv[7]);
        auto code = std::make_shared<ldpc_code>(H, std::string(""));
        ldpc_decoder d(code, dp);
        FILE *fi = fopen(argv[6], "rb"), *f = fopen(argv[8], "wb");
        if (!fi || !f) { fprintf(stderr, "cannot open files\n"); return 2; }
        int32_t hdr[5] = {code->nc(), code->mc(), code->nnz(), nframes, 0};
        put(f, hdr, 5);
        std::vector<double> in(code->nc());
        for (int t = 0; t < nframes; ++t)
        {
            if (fread(in.data(), sizeof(double), in.size(), fi) != in.size()) { fprintf(stderr, "short read\n"); return 2; }
            d.set_llr_in(in);
            int32_t it = d.decode();
            put(f, d.llr_out().data(), in.size()); put_bits(f, d.estimate()); put(f, &it, 1);
        }
        fclose(fi); fclose(f);
        return 0;
    }
    fprintf(stderr, "bad arguments\n");
    return 2;
}

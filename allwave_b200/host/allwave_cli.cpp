// allwave_cli.cpp -- minimal driver with allwave's CLI surface for the alignment path
// (src/main.rs:32-80): -i FASTA [-o PAF] [-s scores] [-p strategy] [-t N] [--wfa-orientation]
// [--no-progress] [--gpu D].  FASTA parsing and the PAF writer stay on the host; everything
// between the pair list and the PAF text runs on the GPU through liballwave_cuda.so.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <memory>

#include "allwave.hpp"

int main(int argc, char** argv) {
    std::string input, output, scores = "0,5,8,2,24,1", spars = "giant:0.99";
    bool wfa_orientation = false, progress = true;
    int device = 0, gpus = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](const char* name) -> std::string {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "error: %s needs a value\n", name);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") input = need("-i");
        else if (a == "-o" || a == "--output") output = need("-o");
        else if (a == "-s" || a == "--scores") scores = need("-s");
        else if (a == "-p" || a == "--sparsification") spars = need("-p");
        else if (a == "-t" || a == "--threads") (void)need("-t");  // host threads are irrelevant: pairs run on the GPU
        else if (a == "--gpu") device = std::atoi(need("--gpu").c_str());
        else if (a == "--gpus") gpus = std::max(1, std::atoi(need("--gpus").c_str()));  // devices device .. device+gpus-1, pairs sharded by predicted cost
        else if (a == "--wfa-orientation") wfa_orientation = true;
        else if (a == "--no-progress") progress = false;
        else {
            std::fprintf(stderr, "usage: allwave -i FASTA [-o PAF] [-s scores] [-p none|auto|random:f|giant:p|tree:n:f:r[:k]] [--wfa-orientation] [--gpu D] [--gpus N]\n");
            return 2;
        }
    }
    if (input.empty()) {
        std::fprintf(stderr, "error: -i/--input is required\n");
        return 2;
    }
    try {
        using namespace allwave;
        const AlignmentParams params = parse_scores(scores);
        const SparsificationStrategy sp = parse_sparsification(spars);
        const std::vector<Sequence> seqs = read_fasta(input);
        Context ctx(device);
        ctx.load(seqs);
        std::vector<std::unique_ptr<Context>> more;
        std::vector<Context*> others;
        for (int g = 1; g < gpus; ++g) {
            more.emplace_back(new Context(device + g));
            more.back()->load(seqs);
            others.push_back(more.back().get());
        }
        AllPairIterator it(ctx, seqs, params, true, !wfa_orientation, sp);
        FILE* out = output.empty() ? stdout : std::fopen(output.c_str(), "w");
        if (!out) throw std::runtime_error("cannot open " + output);
        const auto t0 = std::chrono::steady_clock::now();
        size_t done = 0;
        it.for_each_paf(
            [&](const char* line, size_t len) {
                std::fwrite(line, 1, len, out);
                std::fputc('\n', out);
                ++done;
            },
            others);
        if (out != stdout) std::fclose(out);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (progress) std::fprintf(stderr, "[%.1fs] %zu/%zu (100.0%%) %.1f alignments/sec - Complete!\n", dt, done, it.pair_count(), done / std::max(dt, 1e-9));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}

// run_queries_b200.cpp — the reference's count / locate benchmark driver (benchmark/indexing_count/src/
// run_queries_sdsl.cpp, benchmark/indexing_locate) over this engine: same command line, same pattern-file format
// (Pizza&Chili genpatterns: "# number=N length=M file=F forbidden=...\n" followed by N x M bytes on stdin), same
// "# key = value" report on stderr — but the N patterns are ONE batched call instead of N scalar calls.
//
//   run_queries_b200 <index file> <C|L> [V] [--plain] [--dens D]   < patterns
//
// <index file> is a file the REFERENCE wrote with store_to_file: by default its benchmark index FM_HUFF =
// csa_wt<wt_huff<bit_vector, rank_support_v5<>, select_support_scan<>, select_support_scan<0>>, 1<<20, 1<<20>
// (benchmark/indexing_count/index.config:8); --plain: csa_wt<wt_huff<>> with the default supports; --dens: t_dens.
// Times are wall-clock (the reference reports rusage CPU time, which does not see the GPU).
//
//   g++ -std=c++17 -O2 tools/run_queries_b200.cpp -I. -Lsdsl-lite_b200 -lsdslgpu -Wl,-rpath,$PWD/sdsl-lite_b200 -o run_queries_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../sdsl-lite_b200/include/sdsl_b200.hpp"

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// run_queries_sdsl.cpp:228-254
static void pfile_info(unsigned long * length, unsigned long * numpatt)
{
    char origfilename[257];
    if (std::fscanf(stdin, "# number=%lu length=%lu file=%256s forbidden=", numpatt, length, origfilename) != 3)
    {
        std::fprintf(stderr, "Error: Patterns file header not correct\n");
        std::exit(1);
    }
    std::fprintf(stderr, "# pat_cnt = %lu\n# pat_length = %lu\n# forbidden_chars = ", *numpatt, *length);
    for (int c = std::fgetc(stdin); c != EOF && c != 0 && c != '\n'; c = std::fgetc(stdin))
        std::fprintf(stderr, "%d", c);
    std::fprintf(stderr, "\n");
}

int main(int argc, char ** argv)
{
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s <index file> <C|L> [V] [--plain] [--dens D] < patterns\n", argv[0]);
        return 1;
    }
    bool verbose = false, plain = false;
    uint32_t dens = 1u << 20;
    for (int a = 3; a < argc; ++a)
    {
        if (!std::strcmp(argv[a], "V"))
            verbose = true;
        else if (!std::strcmp(argv[a], "--plain"))
        {
            plain = true;
            dens = 32;
        }
        else if (!std::strcmp(argv[a], "--dens") && a + 1 < argc)
            dens = (uint32_t)std::strtoul(argv[++a], nullptr, 10);
    }
    char const query = argv[2][0];
    sdsl_b200::csa_wt csa;
    std::fprintf(stderr, "# File = %s\n# program = sdsl-lite_b200\n", argv[1]);
    double t0 = now();
    {
        std::ifstream in(argv[1], std::ios::binary);
        if (!in)
        {
            std::fprintf(stderr, "Error: cannot open %s\n", argv[1]);
            return 1;
        }
        csa.load(in, dens, plain ? SDSLGPU_F_DEFAULT : SDSLGPU_F_V5_SCAN);
    }
    double const load_time = now() - t0;
    uint64_t device_bytes = 0;
    sdslgpu_device_bytes(csa.image(), &device_bytes);
    std::fprintf(stderr, "# Load_index_time_in_sec = %.2f\n# text_size = %llu\n# Index_size_in_bytes = %llu\n", load_time,
                 (unsigned long long)(csa.size() - 1), (unsigned long long)device_bytes);
    unsigned long length = 0, numpatt = 0;
    pfile_info(&length, &numpatt);
    std::vector<uint8_t> pats((size_t)length * numpatt);
    if (std::fread(pats.data(), 1, pats.size(), stdin) != pats.size())
    {
        std::fprintf(stderr, "Error: cannot read patterns file\n");
        return 1;
    }
    std::vector<uint64_t> off(numpatt + 1);
    for (unsigned long k = 0; k <= numpatt; ++k)
        off[k] = (uint64_t)k * length;
    if (query == 'C')
    {
        std::vector<uint64_t> cnt(numpatt);
        t0 = now();
        sdsl_b200::check(sdslgpu_fm_count(csa.image(), pats.data(), off.data(), numpatt, cnt.data(), nullptr, nullptr), "count");
        double const t = now() - t0;
        unsigned long long total = 0;
        for (unsigned long k = 0; k < numpatt; ++k)
        {
            total += cnt[k];
            if (verbose)
            { // run_queries_sdsl.cpp:154-159
                std::fputc('C', stdout);
                std::fwrite(&length, sizeof(length), 1, stdout);
                std::fwrite(pats.data() + off[k], 1, length, stdout);
                unsigned long numocc = cnt[k];
                std::fwrite(&numocc, sizeof(numocc), 1, stdout);
            }
        }
        std::fprintf(stderr, "# Total_Num_occs_found = %llu\n# Count_time_in_milli_sec = %.4f\n# Count_time/Pattern_chars = %.6f\n", total, t * 1000,
                     t * 1000 / ((double)length * numpatt));
        std::fprintf(stderr, "# Count_time/Num_patterns = %.6f\n\n# (Load_time+Count_time)/Num_patterns = %.4f\n\n", t * 1000 / numpatt,
                     (load_time + t) * 1000 / numpatt);
    }
    else if (query == 'L')
    {
        std::vector<uint64_t> occ_off(numpatt + 1), occ;
        uint64_t total = 0;
        t0 = now();
        sdsl_b200::check(sdslgpu_fm_locate(csa.image(), pats.data(), off.data(), numpatt, occ_off.data(), nullptr, 0, &total, nullptr), "locate");
        occ.resize(total);
        if (total)
            sdsl_b200::check(sdslgpu_fm_locate(csa.image(), pats.data(), off.data(), numpatt, occ_off.data(), occ.data(), total, &total, nullptr), "locate");
        double const t = now() - t0;
        std::fprintf(stderr, "# Total_Num_occs_found = %llu\n# Locate_time_in_milli_sec = %.4f\n# Locate_time/Num_occs = %.6f\n\n",
                     (unsigned long long)total, t * 1000, total ? t * 1000 / total : 0.0);
    }
    else
    {
        std::fprintf(stderr, "Unknown query type %c (C = count, L = locate)\n", query);
        return 1;
    }
    return 0;
}

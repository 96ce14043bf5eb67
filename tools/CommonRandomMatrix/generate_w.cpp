// CommonRandomMatrix — watermark generator (reference: CommonRandomMatrix/main.cpp,
// samples/make_w.bat).  Writes rows*cols N(0,1) floats, raw little-endian float32, row-major — the file format
// Watermark::loadRandomMatrix expects (Watermark.cpp:62-75).
//
//   CommonRandomMatrix <rows> <cols> <seed> <output_file> [--legacy|--current]
//
// --legacy (default) reproduces the stream of the .dat files the reference COMMITS (samples/w_512.dat, w_480p.dat,
// w_720p.dat, seed 28390211): std::mt19937_64 + std::normal_distribution<float>, single-threaded, with each
// consecutive pair swapped (the MSVC runtime that built them returns the polar pair in the opposite order to
// libstdc++).  Values agree with the committed files to <= 1 ulp (MSVC evaluates sqrt(-2 log s / s) in double).
// --current follows the generator source as it stands today (CommonRandomMatrix/main.cpp:37-51): 32-bit mt19937,
// every OpenMP thread seeded identically over its own chunk — which does NOT reproduce the committed files.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <thread>
#include <vector>

int main(int argc, char* argv[])
{
    if (argc < 5 || argc > 6) {
        std::cerr << "Usage: " << argv[0] << " <rows> <cols> <seed> <output_file> [--legacy|--current]\n";
        return EXIT_FAILURE;
    }
    const int rows = std::stoi(argv[1]), cols = std::stoi(argv[2]);
    const unsigned long long seed = std::stoull(argv[3]);
    const std::string filename = argv[4];
    const bool current = argc == 6 && std::strcmp(argv[5], "--current") == 0;
    if (rows <= 0 || cols <= 0 || rows >= 32768 || cols >= 32768) {  // CommonRandomMatrix/main.cpp:28
        std::cerr << "Rows and columns must be positive integers less than or equal to 32768.\n";
        return EXIT_FAILURE;
    }
    const size_t n = (size_t)rows * cols;
    std::vector<float> w(n);
    if (!current) {
        std::mt19937_64 g(seed);
        std::normal_distribution<float> d(0.0f, 1.0f);
        for (size_t i = 0; i < n; i += 2) {
            const float a = d(g), b = d(g);
            w[i] = b;
            if (i + 1 < n) w[i + 1] = a;
        }
    } else {
        const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        const size_t chunk = n / nt;
        for (unsigned t = 0; t < nt; t++) {  // same values as the OpenMP version: each chunk restarts the same generator
            std::mt19937 g((unsigned)seed);
            std::normal_distribution<float> d(0.0f, 1.0f);
            const size_t b = t * chunk, e = t == nt - 1 ? n : b + chunk;
            for (size_t i = b; i < e; i++) w[i] = d(g);
        }
    }
    std::ofstream out(filename, std::ios::binary);
    if (!out) { std::cerr << "Error: Unable to open file " << filename << " for writing.\n"; return EXIT_FAILURE; }
    out.write(reinterpret_cast<const char*>(w.data()), (std::streamsize)(n * sizeof(float)));
    if (!out) { std::cerr << "Error: Failed to write data to " << filename << ".\n"; return EXIT_FAILURE; }
    std::cout << "Successfully wrote " << n << " random floats to " << filename << ".\n";
    return EXIT_SUCCESS;
}

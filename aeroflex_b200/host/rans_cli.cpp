// rans_cli.cpp -- the RANS leg of AeroFLEX's CLI mode (reference: src/aeroflex/src/app.cpp:168-176, 829-854):
//   rans_cli -i conf.ini [-m mesh_dir] [-a airfoil] [--math strict|fast] [--device d] [--shard i/n] [--vtu dir/] [-q]
// --shard i/n solves the i-th of n contiguous chunks of the alpha list (BASELINE config 5: a 64-angle polar sweep is
// n independent warm-start chains, one per GPU, no communication; scripts/polar_sweep.sh launches them).
// reads the [rans-*] sections of an AeroFLEX conf.ini, runs rans.compute_alphas() and rans.solve_airfoil() on the GPU
// and prints the polar table (alpha, CL, CD, CM) that the VLM viscous correction would consume.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "rans_b200/rans.h"

int main(int argc, char** argv) {
    std::string conf, mesh_dir = "../../../../examples/rans/", airfoil = "naca0012q", math;
    bool quiet = false;
    std::string vtu_dir;
    int shard_i = 0, shard_n = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-i" && i + 1 < argc) conf = argv[++i];
        else if (a == "-m" && i + 1 < argc) mesh_dir = argv[++i];
        else if (a == "-a" && i + 1 < argc) airfoil = argv[++i];
        else if (a == "--math" && i + 1 < argc) math = argv[++i];
        else if (a == "-q") quiet = true;
        else if (a == "--vtu" && i + 1 < argc) vtu_dir = argv[++i];  // write <airfoil>_<alpha>.vtu per angle (rans.h:103) into this directory
        else if (a == "--device" && i + 1 < argc) rans::default_device() = std::atoi(argv[++i]);
        else if (a == "--shard" && i + 1 < argc) { if (std::sscanf(argv[++i], "%d/%d", &shard_i, &shard_n) != 2 || shard_n < 1 || shard_i < 0 || shard_i >= shard_n) { std::fprintf(stderr, "bad --shard\n"); return 2; } }
    }
    if (conf.empty()) { std::fprintf(stderr, "usage: %s -i conf.ini [-m mesh_dir] [-a airfoil] [--math strict|fast] [-q]\n", argv[0]); return 2; }
    if (!math.empty()) setenv("AFX_MATH", math.c_str(), 1);
    try {
        std::cout << "CLI mode" << std::endl;
        GUIHandler gui;
        rans::Rans rans(gui);
        tiny::config io;
        if (!io.read(conf)) return 1;
        rans.settings.import_config_file(io);
        rans.mesh_dir = mesh_dir;
        rans.verbose = !quiet;
        if (!vtu_dir.empty()) { rans.save_vtu = true; rans.vtu_dir = vtu_dir; }
        rans.compute_alphas();
        database::airfoil db;
        {  // this process's contiguous chunk of the alpha list
            const size_t n = rans.alphas.size(), lo = n * (size_t)shard_i / (size_t)shard_n, hi = n * (size_t)(shard_i + 1) / (size_t)shard_n;
            db.alpha.assign(rans.alphas.begin() + (long)lo, rans.alphas.begin() + (long)hi);
        }
        if (db.alpha.empty()) { std::printf("# empty shard\n"); return 0; }
        rans.solve_airfoil(airfoil, db);
        for (auto txt = gui.msg.pop(); txt.has_value(); txt = gui.msg.pop()) std::cout << txt.value() << std::endl;
        std::printf("# alpha_deg CL CD CM\n");
        for (size_t i = 0; i < db.cl.size(); ++i) std::printf("POLAR %.17g %.17g %.17g %.17g\n", db.alpha[i], db.cl[i], db.cd[i], db.cmy[i]);
        std::printf("# outer iterations recorded: %d\n", (int)rans.iters);
    } catch (std::exception& e) {
        std::cout << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// warpii_gpu: command-line front end of the GPU path, with the reference's usage (warpii.cc:31-118):
//
//   warpii_gpu [options] <input_file>      ("-" reads the input from stdin)
//   warpii_gpu --help | -h
//
// Options: --setup-only (stop after setup), --enable-fpe (trap host floating point exceptions while the input is
// evaluated; on the device an unphysical state stops the run with an error instead), --device N.
// The working directory follows the input's WorkDir format (%A__%I) and receives the solution_<n>.vtu frames when write_output is set.
#include <fenv.h>
#include <sys/stat.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "five_moment_app.hpp"

namespace {
void print_help(bool to_err) {
    (to_err ? std::cerr : std::cout) << R"(
warpii_gpu: the FiveMoment application of WarpII on one B200.

Usage:
  warpii_gpu [options] <input_file>
  warpii_gpu --help | -h

Options:
  --setup-only: only perform the setup() phase of the simulation.
  --enable-fpe: trap floating point exceptions on the host (initial and boundary condition evaluation).
  --device N:   CUDA device ordinal (default 0).
)";
}
}  // namespace

int main(int argc, char** argv) {
    bool help = false, fpe = false, setup_only = false;
    int device = 0;
    std::string input_name;
    for (int i = 1; i < argc; i++) {
        const std::string arg = argv[i];
        if (arg == "--help" || arg == "-h") help = true;
        else if (arg == "--enable-fpe") fpe = true;
        else if (arg == "--setup-only") setup_only = true;
        else if (arg == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else input_name = arg;
    }
    if (help) {
        print_help(false);
        return 0;
    }
    if (input_name.empty()) {
        std::cout << "Error: no input source was requested." << std::endl;
        print_help(true);
        return 1;
    }
    std::stringstream text;
    if (input_name == "-") {
        text << std::cin.rdbuf();
    } else {
        std::ifstream file(input_name);
        if (!file.is_open()) {
            std::cerr << "Could not open requested input file <" << input_name << "> for reading." << std::endl;
            print_help(true);
            return 1;
        }
        text << file.rdbuf();
    }
    if (fpe) feenableexcept(FE_DIVBYZERO | FE_INVALID | FE_OVERFLOW);
    try {
        auto app = warpii_b200::FiveMomentGpuApp::create_from_input(text.str(), 0, 1, device);
        // remove_file_extension + format_workdir (warpii.cc:198-219)
        std::string stem = input_name == "-" ? "STDIN" : input_name;
        if (input_name != "-") {
            const size_t slash = stem.find_last_of('/');
            const size_t dot = stem.find_last_of('.');
            if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) stem.erase(dot);
        }
        const std::string workdir = app->format_workdir(stem);
        struct stat info;
        if (stat(workdir.c_str(), &info) != 0) {
            if (mkdir(workdir.c_str(), 0755) != 0) {
                std::cerr << "mkdir() error: " << std::strerror(errno) << std::endl;
                return 1;
            }
            std::cout << "Directory created: " << workdir << std::endl;
        } else if (!(info.st_mode & S_IFDIR)) {
            std::cerr << "Error: " << workdir << " is not a directory." << std::endl;
            return 1;
        }
        app->set_output_dir(workdir);
        app->set_frame_callback([](unsigned frame, double t) { std::cout << "frame " << frame << "  t = " << t << std::endl; });
        std::cout << "Setting up" << std::endl;
        app->setup();
        if (setup_only) return 0;
        app->run();
        std::cout << "steps = " << app->get_solver().steps_taken() << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// warpii_gpu: command-line front end of the GPU path, with the reference's usage (warpii.cc:31-118):
//
//   warpii_gpu [options] <input_file>      ("-" reads the input from stdin)
//   warpii_gpu --help | -h
//
// Options: --setup-only (stop after setup), --enable-fpe (trap host floating point exceptions while the input is
// evaluated; on the device an unphysical state stops the run with an error instead), --device N, --gpus N (one process per
// GPU; where the reference is started under mpirun, this launcher forks the ranks itself).
// The working directory follows the input's WorkDir format (%A__%I) and receives the solution_<n>.vtu frames when write_output is set.
//
// Header form so that an extension example links the same front end: warpii_cli_main(argc, argv, std::make_shared<MyExtension>())
// is the GPU path's `Warpii::create_from_cli(argc, argv, ext).run()` (examples/five-moment/forward_facing_step/main.cc).
#pragma once
#include <fenv.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <array>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "five_moment_app.hpp"

namespace {
inline void print_help(bool to_err) {
    (to_err ? std::cerr : std::cout) << R"(
warpii_gpu: the FiveMoment application of WarpII on one B200.

Usage:
  warpii_gpu [options] <input_file>
  warpii_gpu --help | -h

Options:
  --setup-only: only perform the setup() phase of the simulation.
  --enable-fpe: trap floating point exceptions on the host (initial and boundary condition evaluation).
  --device N:   CUDA device ordinal (default 0); with --gpus the first of N consecutive devices.
  --gpus N:     shard the elements over N GPUs of this box (one process per GPU, NCCL halo exchange);
                frames are written as solution_<n>.rank<r>.vtu with a solution_<n>.pvtu index.
)";
}
}  // namespace

// One rank of a run: rank 0 of a sharded run creates the NCCL id and hands it to the launcher through id_out_fd; the other
// ranks read it from id_in_fd.  A single-GPU run passes -1 for both.
inline int run_rank(const std::string& text, const std::string& workdir, int rank, int n_ranks, int device, bool setup_only,
                    int id_out_fd, int id_in_fd, std::shared_ptr<warpii_b200::GridExtension> ext = nullptr) {
    try {
        auto app = warpii_b200::FiveMomentGpuApp::create_from_input(text, rank, n_ranks, device, ext);
        app->set_output_dir(workdir);
        if (rank == 0)
            app->set_frame_callback([](unsigned frame, double t) { std::cout << "frame " << frame << "  t = " << t << std::endl; });
        if (rank == 0) std::cout << "Setting up" << std::endl;
        char id[WARPII_GPU_NCCL_ID_BYTES];
        if (n_ranks > 1) {   // exchange the id before any device work so that a failing rank cannot leave the others waiting
            if (rank == 0) {
                if (warpii_gpu_nccl_unique_id(id) != 0) throw std::runtime_error(warpii_gpu_last_error());
                if (write(id_out_fd, id, sizeof id) != (ssize_t)sizeof id) throw std::runtime_error("cannot hand the NCCL id to the launcher");
            } else if (read(id_in_fd, id, sizeof id) != (ssize_t)sizeof id) {
                throw std::runtime_error("did not receive the NCCL id");
            }
        }
        app->setup();
        if (n_ranks > 1) app->attach_comm(id);
        if (setup_only) return 0;
        app->run();
        if (rank == 0) std::cout << "steps = " << app->get_solver().steps_taken() << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "Error" << (n_ranks > 1 ? " (rank " + std::to_string(rank) + ")" : std::string()) << ": " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// Warpii::create_from_cli + run (warpii.cc:31-118, 126-196); ext as in the reference's extension examples
inline int warpii_cli_main(int argc, char** argv, std::shared_ptr<warpii_b200::GridExtension> ext = nullptr) {
    bool help = false, fpe = false, setup_only = false;
    int device = 0, gpus = 1;
    std::string input_name;
    for (int i = 1; i < argc; i++) {
        const std::string arg = argv[i];
        if (arg == "--help" || arg == "-h") help = true;
        else if (arg == "--enable-fpe") fpe = true;
        else if (arg == "--setup-only") setup_only = true;
        else if (arg == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (arg == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
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
    if (gpus < 1) {
        std::cerr << "Error: --gpus needs a positive number." << std::endl;
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
    std::string workdir;
    try {
        // parse once in the launcher (no device involved): input errors surface before anything is forked
        auto probe = warpii_b200::FiveMomentGpuApp::create_from_input(text.str(), 0, gpus, device, ext);
        // remove_file_extension + format_workdir (warpii.cc:198-219)
        const std::string stem = input_name == "-" ? "STDIN" : warpii_b200::FiveMomentGpuApp::remove_file_extension(input_name);
        workdir = probe->format_workdir(stem);
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
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
    if (gpus == 1) return run_rank(text.str(), workdir, 0, 1, device, setup_only, -1, -1, ext);

    // One process per GPU (devices device .. device+gpus-1).  The launcher never touches CUDA: it forks the ranks, relays
    // the 128-byte NCCL id from rank 0 to the others over pipes and collects the exit codes.
    std::cout.flush();
    int up[2];
    if (pipe(up) != 0) { std::perror("pipe"); return 1; }
    std::vector<std::array<int, 2>> down(gpus);
    std::vector<pid_t> pids(gpus);
    for (int r = 0; r < gpus; r++) {
        if (r > 0 && pipe(down[r].data()) != 0) { std::perror("pipe"); return 1; }
        pids[r] = fork();
        if (pids[r] < 0) { std::perror("fork"); return 1; }
        if (pids[r] == 0) {
            // Keep only the ends this rank uses.  Every other inherited end must go: a child that kept the write end of `up`
            // (or of an earlier rank's `down` pipe) would keep those pipes open after rank 0 has died without sending the id,
            // and the launcher's read -- and with it every other rank's read -- would never see end-of-file.
            close(up[0]);
            if (r != 0) close(up[1]);
            for (int k = 1; k < r; k++) close(down[k][1]);     // (their read ends were closed by the launcher before this fork)
            if (r > 0) close(down[r][1]);
            const int rc = run_rank(text.str(), workdir, r, gpus, device + r, setup_only, r == 0 ? up[1] : -1, r > 0 ? down[r][0] : -1, ext);
            std::cout.flush();
            _exit(rc);
        }
        if (r > 0) close(down[r][0]);
    }
    close(up[1]);
    char id[WARPII_GPU_NCCL_ID_BYTES];
    const bool have_id = read(up[0], id, sizeof id) == (ssize_t)sizeof id;
    for (int r = 1; r < gpus; r++) {
        if (have_id && write(down[r][1], id, sizeof id) != (ssize_t)sizeof id) std::perror("write");
        close(down[r][1]);   // without an id the readers see end-of-file and stop with an error
    }
    int worst = have_id ? 0 : 1;
    for (int r = 0; r < gpus; r++) {
        int status = 0;
        waitpid(pids[r], &status, 0);
        const int rc = WIFEXITED(status) ? WEXITSTATUS(status) : 1;
        if (rc > worst) worst = rc;
    }
    return worst;
}

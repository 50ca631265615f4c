// warpii_gpu: command-line front end of the GPU path (see warpii_cli.hpp), no extension.
#include "warpii_cli.hpp"

int main(int argc, char** argv) { return warpii_cli_main(argc, argv); }

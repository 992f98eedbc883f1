// upcgen command line, same flags as the reference's main.cpp:77-99:
//   ./upcgen [-debug N] [-nthreads N] [-parfile F] [-device D] [-ngpus G] [-h]
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "UpcGenerator.h"

int main(int argc, char** argv)
{
  auto* upcGenerator = new UpcGenerator();
  for (int i = 1; i < argc; i++) {
    if (std::strcmp(argv[i], "-h") == 0) {
      std::cout << "Available options:\n"
                << "  -debug N      debug level\n"
                << "  -nthreads N   accepted for compatibility (the table stage runs on the GPU)\n"
                << "  -parfile F    parameters file (default parameters.in)\n"
                << "  -device D     CUDA device (default 0)\n"
                << "  -ngpus G      number of GPUs (devices D .. D+G-1): the (y, m) grid is sharded by m rows, the table\n"
                << "                all-gathered over NCCL, the events split by candidate ranges (default 1)\n";
      return 0;
    }
    if (std::strcmp(argv[i], "-debug") == 0 && i + 1 < argc) upcGenerator->setDebugLevel(std::atoi(argv[++i]));
    else if (std::strcmp(argv[i], "-nthreads") == 0 && i + 1 < argc) upcGenerator->setNumThreads(std::atoi(argv[++i]));
    else if (std::strcmp(argv[i], "-parfile") == 0 && i + 1 < argc) upcGenerator->setParFile(argv[++i]);
    else if (std::strcmp(argv[i], "-device") == 0 && i + 1 < argc) upcGenerator->setDevice(std::atoi(argv[++i]));
    else if (std::strcmp(argv[i], "-ngpus") == 0 && i + 1 < argc) upcGenerator->setNumGpus(std::atoi(argv[++i]));
  }
  upcGenerator->configGeneratorFromFile();
  upcGenerator->init();
  upcGenerator->generateEvents();
  std::cout << "total cross section [mb]: " << upcGenerator->totNuclX() << "  fiducial [mb]: " << upcGenerator->fidNuclX()
            << std::endl;
  delete upcGenerator;
  return 0;
}

// parse_check -- parameters.in through the DROP-IN's parser (upcgen_b200/host/UpcGenerator.cpp): prints the parameter
// block as the same "KEY value" lines oracle/refshim/gen_capi.cpp prints for the reference's parser.  No GPU needed:
// nothing is initialised.
#include <iostream>

#include "UpcCrossSection.h"
#include "UpcSampler.h"
#define private public
#include "UpcGenerator.h"
#undef private
#include "../../oracle/refshim/param_dump.h"

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  UpcGenerator g;
  g.setDebugLevel(0);
  g.setParFile(argv[1]);
  g.configGeneratorFromFile();
  std::cout << upc_param_dump(g);
  return 0;
}

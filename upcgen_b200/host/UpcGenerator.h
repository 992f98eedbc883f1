// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// UpcGenerator -- orchestration class with the reference's public interface
// (include/UpcGenerator.h:58-128): parameters.in parsing, init(), computeNuclXsection(),
// generateEvent()/generateEvents(), HepMC3-ASCII output.  The tables come from UpcCrossSection
// (GPU); events are produced on the GPU in blocks of candidates (upcgpu_generate) and handed out
// one at a time through generateEvent().
#pragma once

#include <charconv>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

#include "UpcCompat.h"
#include "UpcCrossSection.h"
#include "UpcSampler.h"

class UpcGenerator
{
 public:
  UpcGenerator();
  ~UpcGenerator();

  // process-specific parameters
  int procID{0};
  double aLep{0};

  // simulation parameters
  bool doPtCut{false};
  double minPt{0};
  bool doEtaCut{false};
  double minEta{0};
  double maxEta{0};
  bool usePolarizedCS{false};
  long int seed{0};
  long int nEvents{1000};

  void setParameterValue(const std::string& parameter, const std::string& parValue);
  static int debug;
  void configGeneratorFromFile();
  void init();
  void setLumiFileDirectory(std::string directory) { nucProcessCS->setLumiFileDirectory(directory); };
  void setDebugLevel(int level) { debug = level; }
  void setNumThreads(int n) { numThreads = n; }
  void setParFile(std::string parfilename) { parFileName = parfilename; }
  void setCollisionSystem(float sqrts, int nucl_z, int nucl_a);
  void setSeed(int seedIn) { seed = seedIn; PLOG_INFO << "<seed> = " << seed; }
  void printParameters();
  void computeNuclXsection();
  double totNuclX() { return totCS; }
  double fidNuclX() { return fidCS; }
  long int generateEvent(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                         std::vector<TLorentzVector>& particles);
  const std::vector<TParticle>& getParticles() const { return genParticles; };
  void generateEvents();

  // additions of the GPU build
  void setDevice(int dev) { nucProcessCS->device = dev; }
  void setNumGpus(int n) { nucProcessCS->numGpus = n < 1 ? 1 : n; }
  UpcCrossSection* crossSection() { return nucProcessCS; }

 private:
  UpcCrossSection* nucProcessCS{nullptr};
  std::string parFileName{"parameters.in"};
  bool useROOTOut{true};
  bool useHepMCOut{false};
  double totCS;
  double fidCS;
  bool ignoreCSZ;
  std::vector<TParticle> genParticles;
  std::vector<std::vector<double>> nucCSYM;
  std::vector<std::vector<double>> nucTargRatioCSYM; // for VM production
  std::vector<std::vector<double>> polCSRatio;
  std::vector<double> binEdgesM, binEdgesZ, binEdgesY;

  int pythiaVersion{-1};
  bool isPythiaUsed{false};
  bool doFSR{false};
  bool doDecays{false};
  int numThreads{1};
  bool isPairProduction{false};
  bool isSingleProduction{false};
  bool isPairProductionVM{false};
  // vector-meson events are generated on the host (one (y) draw, rejection sampling of the pomeron pT and of the decay
  // angle: src/UpcGenerator.cpp:425-472, src/UpcCrossSection.cpp:1076-1104)
  UpcSampler2D* samplerCsYM{nullptr};
  long int generateEventVM(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                           std::vector<TLorentzVector>& particles);
  void twoPartDecayVM(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                      std::vector<TLorentzVector>& particles, int id);
  bool checkKinCuts(std::vector<TLorentzVector>& particles);

  // a very simple HepMC writer: particles only, no vertex information (as the reference's)
  class WriterHepMC
  {
   public:
    explicit WriterHepMC(const std::string& fname) { openFile(fname); }
    ~WriterHepMC() { closeFile(); }
    std::ofstream outfile;
    void openFile(const std::string& fname)
    {
      outfile.open(fname);
      outfile << "HepMC::Version 3.02.04\n" << "HepMC::Asciiv3-START_EVENT_LISTING\n";
    }
    void closeFile()
    {
      outfile << "HepMC::Asciiv3-END_EVENT_LISTING\n";
      outfile.close();
    }
    // The records are the reference's (include/UpcGenerator.h:214-241: stream insertion with setprecision(9), i.e.
    // printf's %.9g); they are formatted with std::to_chars -- defined to give printf's characters -- into a line
    // buffer and handed to the stream in one write: three times the pace of nine operator<< calls per particle,
    // which was the larger part of writing events.hepmc.
    void writeEventInfo(long int eventID, int nParticles, int nVertices = 0)
    {
      char line[96];
      char* q = line;
      *q++ = 'E'; *q++ = ' ';
      q = putInt(q, eventID); *q++ = ' ';
      q = putInt(q, nVertices); *q++ = ' ';
      q = putInt(q, nParticles);
      std::memcpy(q, "\nU GEV MM\n", 10); q += 10;
      outfile.write(line, q - line);
    }
    void writeParticleInfo(int id, int motherID, int pdg, double px, double py, double pz, double e, double m, int status)
    {
      char line[256];
      char* q = line;
      *q++ = 'P'; *q++ = ' ';
      q = putInt(q, id); *q++ = ' ';
      q = putInt(q, motherID); *q++ = ' ';
      q = putInt(q, pdg); *q++ = ' ';
      const double v[5] = {px, py, pz, e, m};
      for (double x : v) { q = std::to_chars(q, q + 32, x, std::chars_format::general, 9).ptr; *q++ = ' '; }
      q = putInt(q, status);
      *q++ = '\n';
      outfile << std::setprecision(9);  // the stream is left as the reference's writer leaves it
      outfile.write(line, q - line);
    }

    static char* putInt(char* q, long v) { return std::to_chars(q, q + 24, v).ptr; }
  };
  WriterHepMC* writerHepMC{nullptr};
  std::vector<std::vector<double>> treeCols;  // the nine branches of the tree "particles" (events.root)
  void writeEvent(long int evt, const std::vector<int>& pdgs, const std::vector<int>& statuses,
                  const std::vector<int>& mothers, const std::vector<TLorentzVector>& particles);

  // block of candidates generated on the GPU and consumed by generateEvent()
  struct Block {
    std::vector<int> npart, pdg, status, mother;
    std::vector<double> p4;
    size_t pos{0}, n{0};
    int stride{4};  // particle slots per candidate in pdg/status/mother/p4 (upcgpu_generate_packed)
  } block;
  uint64_t nextCandidate{0};
  void refillBlock();
};

// nearest2HammingExhaustiveOMPMEX.cpp -- same contract as PP/mex/nearest2HammingExhaustiveOMPMEX.cpp:18-83
// (the reference's OpenMP variant); on the GPU both names run the same kernel.
#include "nearest2HammingExhaustiveMEX.cpp"

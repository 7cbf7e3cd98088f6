// gp_ss_ak command line (train / test) of the B200 exact-GP path.  The class below keeps the NAME and the three entry points
// of the reference's command object so that code written against it still links; everything numerical it drives lives in
// GP_utils (device-resident, include/gpss.h).
#ifndef GPSS_HOST_GP_SS_AK_H
#define GPSS_HOST_GP_SS_AK_H

#include <armadillo>

#include "Control.h"      // argument cursor, data reader, standardisation
#include "GP_Utils.h"     // the model: objective / gradient / prediction over the C ABI
#include "Kernel.h"       // Kernels, HybKerns, Kern_ExpAnisotropic, Kern_Exponential, Kern_RBF, Kern_Bias
#include "Opt_pars.h"     // BFGS / L-BFGS / SCG drivers

class GP_Cntrl : public Control {
 public:
  GP_Cntrl(int argc, char** argv);

  // usage text for the current mode ("gp", "train" or "test")
  void Help();
  // `train [-k K] [-kn 0|1] [-o OPT] [-# iters] data [model]`: fit the hyper-parameters, write <model>, <model>_Statistics.txt,
  // <model>_predict.txt and the gnuplot script
  void train();
  // `test data model train_data [predictions]`: predictive mean / standard deviation of the rows of `data`
  void test();
};

int main(int argc, char* argv[]);

#endif

// gp_ss_ak command line (train / test) -- the reference's GP_Cntrl surface (gp_ss_ak.h in /root/reference).
#ifndef GPSS_HOST_GP_SS_AK_H
#define GPSS_HOST_GP_SS_AK_H

#include <armadillo>

#include "Control.h"
#include "GP_Utils.h"
#include "Kernel.h"
#include "Opt_pars.h"

int main(int argc, char* argv[]);

class GP_Cntrl : public Control {
 public:
  GP_Cntrl(int argc, char** argv);
  void train();
  void test();
  void Help();
};

#endif

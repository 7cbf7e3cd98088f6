// CPU test harness for the host-only pieces: data reader, symmetric standardisation (train and test mode), the
// <model>_Statistics.txt and train_model files.  No GPU call is made (GP_utils is only used for its (de)serialisation).
//   host_io_check train.txt test.txt theta.txt workdir   ->  dump on stdout, files in workdir
#include <cstdio>
#include <fstream>

#include "../gp_ss_ak.h"

using namespace arma;

static void dump(const char* key, const mat& m)
{
  printf("%s %llu %llu", key, (unsigned long long)m.n_rows, (unsigned long long)m.n_cols);
  for (uword i = 0; i < m.n_elem; i++) printf(" %.17g", m[i]);
  printf("\n");
}

int main(int argc, char** argv)
{
  if (argc < 5) return 2;
  const std::string model = std::string(argv[4]) + "/host_model";
  char* fake[] = {argv[0], 0};
  int mode_train = 0, mode_test = 1;
  bool yscale = true;

  Control ctl(1, fake);
  ctl.setMode("train");
  ctl.setprepM(1);
  int* sz = ctl.readDataSize(argv[1]);
  mat X(sz[0], sz[1]), y(sz[0], 1);
  ctl.readDataFile(X, y, sz, argv[1]);
  dump("X_raw", X);
  dump("y_raw", y);
  ctl.prepareData(X, y, mode_train, yscale, model);
  dump("Xs", X);
  dump("ys", y);
  dump("params", ctl.params);

  Control ctl2(1, fake);
  ctl2.setMode("test");
  ctl2.setprepM(1);
  int* szt = ctl2.readDataSize(argv[2]);
  mat Xt(szt[0], szt[1]), yt(szt[0], 1);
  ctl2.readDataFile(Xt, yt, szt, argv[2]);
  ctl2.prepareData(Xt, yt, mode_test, yscale, model);
  dump("Xt", Xt);
  mat back = Xt, yb = yt;
  ctl2.postData(back, yb, yscale, model);
  dump("Xt_back", back);

  // model file: parameters from theta.txt -> write -> read back
  HybKerns Kerns(X);
  Kern_ExpAnisotropic ke(X);
  Kern_Bias kb(X);
  Kerns.addNewKernel(&ke);
  Kerns.addNewKernel(&kb);
  GP_utils gp(&Kerns, X, y, GP_utils::inf_laplace, GP_utils::likeL_Gaussian, GP_utils::mean_zero, 8, 1, 0, 0);
  mat th(1, gp.getNumPars());
  dump("theta_default", (gp.get_GP_Pars(th), th));
  std::ifstream tf(argv[3]);
  for (uword i = 0; i < th.n_elem; i++) tf >> th[i];
  gp.set_GP_Pars(th);
  writeGPFile(gp, model, "# GP_SS_AK Model File ");
  GP_utils* back_gp = readGpFromFile(model, 0);
  mat th2(1, back_gp->getNumPars());
  back_gp->get_GP_Pars(th2);
  dump("theta_readback", th2);
  printf("readback numData %u inputDim %u outputDim %u kernel %s nkern_params %u\n", back_gp->getNumData(), back_gp->getInpDim(),
         back_gp->getOutDim(), back_gp->KerenlW->getKerName().c_str(), back_gp->KerenlW->getNPars());
  return 0;
}

// Model metadata holders (names and sizes) -- the reference's ModelInfo / Modeling (ModelInf.h:22-179) --
// and its sign() helper, whose sign(0) = -1 matters to the Brent search (ModelInf.h:14-20).
#ifndef GPSS_HOST_MODELINF_H
#define GPSS_HOST_MODELINF_H

#include <cstdlib>
#include <iostream>
#include <string>

#include "StreamInt.h"
#include <armadillo>
using arma::mat;

inline double sign(double val) { return (val <= 0) ? -1.0 : 1.0; }

class ModelInfo {
 public:
  ModelInfo() : numData(0) {}
  explicit ModelInfo(unsigned int nData) : numData(nData) {}
  virtual ~ModelInfo() {}

  std::string getName() const { return modelName; }
  void setName(const std::string name) { modelName = name; }
  std::string getInf() const { return InfName; }
  void setInf(const std::string name) { InfName = name; }
  std::string getlik() const { return likeName; }
  void setlik(const std::string name) { likeName = name; }
  std::string getMean() const { return MeanName; }
  void setMean(const std::string name) { MeanName = name; }

  virtual unsigned int getNumPars() const = 0;
  virtual void ShowKernelPars(std::ostream& os) const = 0;
  virtual unsigned int getNumData() const { return numData; }
  virtual void setNumData(unsigned int val) { numData = val; }
  void ErrorTermination(const std::string error)
  {
    std::cerr << error << std::endl << std::endl;
    std::exit(1);
  }

 private:
  std::string modelName, InfName, likeName, MeanName;
  unsigned int numData;
};

class Modeling : public ModelInfo {
 public:
  Modeling() : ModelInfo(), outputDim(0), inputDim(0), NumMF(0), Numlikf(0), NumCov(0) {}
  Modeling(unsigned int inDim, unsigned int outDim, unsigned int nData)
      : ModelInfo(nData), outputDim(outDim), inputDim(inDim), NumMF(0), Numlikf(0), NumCov(0) {}

  virtual void Calc_Out(mat& yPred, const mat& inData) const = 0;

  void setInpDim(unsigned int dim) { inputDim = dim; }
  unsigned int getInpDim() const { return inputDim; }
  void setOutDim(unsigned int dim) { outputDim = dim; }
  unsigned int getOutDim() const { return outputDim; }
  void setNumMFpar(unsigned int v) { NumMF = v; }
  unsigned int getNumMFpar() const { return NumMF; }
  void setNumCovpar(unsigned int v) { NumCov = v; }
  unsigned int getNumCovpar() const { return NumCov; }
  void setNumlikfpar(unsigned int v) { Numlikf = v; }
  unsigned int getNumlikfpar() const { return Numlikf; }

 private:
  unsigned int outputDim, inputDim, NumMF, Numlikf, NumCov;
};

#endif

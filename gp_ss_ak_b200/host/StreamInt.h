// Text (de)serialisation interface of the train_model file: one `key=value` per line, lines starting with
// '#' are comments.  Mirrors the reference's StreamIntfce (StreamInt.h:46-123): same member names, same
// parsing rule (everything after the first '=' is the value; '#' lines are skipped only here, not by raw
// getline callers -- Kernel.cpp:1285,1318).
#ifndef GPSS_HOST_STREAMINT_H
#define GPSS_HOST_STREAMINT_H

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "DistHost.h"

class StreamIntfce {
 public:
  virtual ~StreamIntfce() {}

  virtual void ToFile_GP_Params(std::ostream& out) const = 0;
  virtual void FromFile_GP_Params(std::istream& in) = 0;
  virtual void StrmOut(std::ostream& out) const { ToFile_GP_Params(out); }
  virtual void StrmIn(std::istream& in) { FromFile_GP_Params(in); }

  // value part of the next non-comment line (StreamInt.h:75-89)
  static std::string ReadStrStrm(std::istream& in, const std::string /*fieldName*/)
  {
    std::string line;
    std::getline(in, line);
    while (line.compare(0, 1, "#") == 0) std::getline(in, line);
    const std::string::size_type eq = line.find("=");
    return line.substr(eq + 1);           // npos + 1 == 0: a line without '=' is returned whole, as in the reference
  }
  static int ReadIntStrm(std::istream& in, const std::string field) { return (int)std::atol(ReadStrStrm(in, field).c_str()); }
  static double ReadDoubleStrm(std::istream& in, const std::string field) { return std::atof(ReadStrStrm(in, field).c_str()); }
  static bool ReadBoolStrm(std::istream& in, const std::string field) { return std::atol(ReadStrStrm(in, field).c_str()) != 0; }

  // comment line first, then the object (StreamInt.h:102-111)
  void WFile(const std::string fileName, const std::string comment = "") const
  {
    std::ofstream out(gpss_host::out_path(fileName).c_str());
    if (!out) {
      std::cout << "The file " << fileName << " is open.\n";
      std::exit(1);
    }
    out << comment << std::endl;
    StrmOut(out);
    out.close();
  }
  void RFile(const std::string fileName)
  {
    std::ifstream in(fileName.c_str());
    if (!in.is_open()) {
      std::cout << "The file could not be read. \n";
      std::exit(1);
    }
    StrmIn(in);
    in.close();
  }
};

#endif

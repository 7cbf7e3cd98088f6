// make_block_model -- writes the test file of BASELINE.json configs[3]: block-model centroids on a regular nx x ny x nz grid over a
// bounding box, one "x<TAB>y<TAB>z<TAB>grade" line per block (the format Control::readDataFile reads, Control.cpp:27-141 of the
// reference).  The grade column is a smooth synthetic function of position: `gp_ss_ak test` wants an observed value per row (it
// reports the MSE and sorts the prediction file by it), a block model has none.
//   make_block_model out.txt nx ny nz lox loy loz hix hiy hiz
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv)
{
  if (argc != 11) { std::fprintf(stderr, "usage: %s out.txt nx ny nz lox loy loz hix hiy hiz\n", argv[0]); return 2; }
  const long n[3] = {std::atol(argv[2]), std::atol(argv[3]), std::atol(argv[4])};
  double lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = std::atof(argv[5 + d]); hi[d] = std::atof(argv[8 + d]); }
  FILE* f = std::fopen(argv[1], "w");
  if (!f) { std::perror(argv[1]); return 1; }
  std::vector<char> buf(1 << 22);
  std::setvbuf(f, buf.data(), _IOFBF, buf.size());
  for (long i = 0; i < n[0]; i++) {
    const double x = lo[0] + (i + 0.5) * (hi[0] - lo[0]) / n[0];
    for (long j = 0; j < n[1]; j++) {
      const double y = lo[1] + (j + 0.5) * (hi[1] - lo[1]) / n[1];
      for (long k = 0; k < n[2]; k++) {
        const double z = lo[2] + (k + 0.5) * (hi[2] - lo[2]) / n[2];
        const double g = 1.0 + 0.45 * std::sin(x / 97.0) * std::cos(y / 131.0) + 0.25 * std::sin((x + y + 3.0 * z) / 211.0);
        std::fprintf(f, "%.10g\t%.10g\t%.10g\t%.8g\n", x, y, z, g);
      }
    }
  }
  std::fclose(f);
  return 0;
}

// DistHost.h -- multi-GPU launch of the host program: one process per GPU, every process runs the SAME deterministic
// train / test flow (the distributed C ABI returns identical results on every rank), rank 0 alone prints and writes files.
// The reference is a single process (no counterpart); this is the host half of DESIGN.md section 7.
//
// Environment (torchrun's names are understood as well, so `python -m torch.distributed.run --no-python ./gp_ss_ak ...`
// works, as does scripts/run_dist_cli.sh):
//   GPSS_WORLD | WORLD_SIZE     number of processes (default 1: everything below is inert)
//   GPSS_RANK  | RANK           this process
//   GPSS_DEVICE | LOCAL_RANK    CUDA device of this process
//   GPSS_RENDEZVOUS_DIR         directory visible to all ranks for the 128-byte NCCL id file (default /tmp)
//   GPSS_JOB | MASTER_PORT      distinguishes concurrent jobs in that directory (default: the launcher's pid)
#ifndef GPSS_DIST_HOST_H
#define GPSS_DIST_HOST_H

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <ctime>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace gpss_host {

inline int env_int(const char* a, const char* b, int dflt)
{
  const char* v = std::getenv(a);
  if (!v && b) v = std::getenv(b);
  return v ? std::atoi(v) : dflt;
}
inline int world() { static const int w = env_int("GPSS_WORLD", "WORLD_SIZE", 1); return w < 1 ? 1 : w; }
inline int rank() { static const int r = env_int("GPSS_RANK", "RANK", 0); return world() > 1 ? r : 0; }
inline int device() { return env_int("GPSS_DEVICE", world() > 1 ? "LOCAL_RANK" : 0, 0); }

// Files are written by rank 0 only; the other ranks run the same code against /dev/null.
inline std::string out_path(const std::string& p) { return rank() == 0 ? p : std::string("/dev/null"); }

// Ranks != 0 print nothing (their stdout is pointed at /dev/null once, at start-up).
inline void silence_other_ranks()
{
  if (rank() == 0) return;
  static std::ofstream devnull("/dev/null");
  std::cout.rdbuf(devnull.rdbuf());
}

// The 128-byte NCCL unique id travels through a file: rank 0 writes <dir>/gpss_id_<job>_<seq> atomically, the others
// poll for it (seq = how many communicators this process has created: train and test create one handle each).
inline std::string id_file(int seq)
{
  const char* dir = std::getenv("GPSS_RENDEZVOUS_DIR");
  const char* job = std::getenv("GPSS_JOB");
  if (!job) job = std::getenv("MASTER_PORT");
  std::string name = std::string(dir ? dir : "/tmp") + "/gpss_id_";
  name += job ? std::string(job) : std::to_string((long)getppid());
  return name + "_" + std::to_string(seq);
}
// A job that died between publish_id and rank 0's remove leaves its file behind, and the key repeats (MASTER_PORT is constant
// under torchrun, pids are recycled).  So: (1) rank 0 removes any leftover of its key at start-up (clear_stale_ids, called before
// anything is published) and creates the file with O_EXCL under a temporary name before renaming it in; (2) readers accept only a
// file written after their own process started (minus the launch skew between ranks) -- a leftover is older than that.
inline time_t process_start() { static const time_t t0 = std::time(nullptr); return t0; }
inline void clear_stale_ids()
{
  process_start();
  if (world() <= 1 || rank() != 0) return;
  for (int seq = 0; seq < 8; seq++) { std::remove(id_file(seq).c_str()); std::remove((id_file(seq) + ".tmp").c_str()); }
}
inline bool publish_id(const std::string& path, const unsigned char id[128])
{
  const std::string tmp = path + ".tmp";
  std::remove(tmp.c_str());
  std::remove(path.c_str());
  const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
  if (fd < 0) return false;
  const bool ok = ::write(fd, id, 128) == 128;
  ::close(fd);
  return ok && std::rename(tmp.c_str(), path.c_str()) == 0;
}
inline bool fetch_id(const std::string& path, unsigned char id[128], int timeout_s = 300, int launch_skew_s = 30)
{
  const time_t oldest = process_start() - launch_skew_s;
  for (int waited_ms = 0; waited_ms < timeout_s * 1000; waited_ms += 20) {
    struct stat sb;
    if (::stat(path.c_str(), &sb) == 0 && sb.st_mtime >= oldest) {
      FILE* f = std::fopen(path.c_str(), "rb");
      if (f) {
        const size_t got = std::fread(id, 1, 128, f);
        std::fclose(f);
        if (got == 128) return true;
      }
    }
    usleep(20000);
  }
  std::fprintf(stderr, "gp_ss_ak: rank %d timed out after %d s waiting for the NCCL id file %s (is rank 0 running? a file older than this "
               "process is ignored as a leftover)\n", rank(), timeout_s, path.c_str());
  return false;
}

}  // namespace gpss_host
#endif

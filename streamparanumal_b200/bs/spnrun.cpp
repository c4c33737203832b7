// spnrun — local rank launcher standing in for `mpirun -np P ./BSn ...` (the reference's run scripts,
// e.g. BS/BS8/runBS8.sh; this image has no MPI).  Forks P copies of the command on this node with
//   SPN_RANK / SPN_SIZE / LOCAL_RANK / SPN_RENDEZVOUS
// set; spn::comm_t::World() builds its socket mesh from them and platform_t binds rank r to GPU r (mod count).
//   usage:  spnrun -np P <program> [args...]
// Exit status: 0 if every rank exits 0, else the first non-zero status; when a rank fails the others are
// terminated (by the pids this launcher created).
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char** argv) {
  int np = 1, i = 1;
  while (i < argc) {
    if ((!strcmp(argv[i], "-np") || !strcmp(argv[i], "-n")) && i + 1 < argc) {
      np = atoi(argv[i + 1]);
      i += 2;
    } else {
      break;
    }
  }
  if (np < 1 || i >= argc) {
    fprintf(stderr, "usage: %s -np P <program> [args...]\n", argv[0]);
    return 2;
  }
  char tmpl[] = "/tmp/spnrun_XXXXXX";
  const char* dir = mkdtemp(tmpl);
  if (!dir) {
    perror("spnrun: mkdtemp");
    return 2;
  }
  std::vector<pid_t> pids;
  for (int r = 0; r < np; ++r) {
    const pid_t pid = fork();
    if (pid < 0) {
      perror("spnrun: fork");
      for (pid_t p : pids) kill(p, SIGTERM);
      return 2;
    }
    if (pid == 0) {
      setenv("SPN_RANK", std::to_string(r).c_str(), 1);
      setenv("SPN_SIZE", std::to_string(np).c_str(), 1);
      setenv("LOCAL_RANK", std::to_string(r).c_str(), 1);
      setenv("SPN_RENDEZVOUS", dir, 1);
      execvp(argv[i], argv + i);
      fprintf(stderr, "spnrun: cannot exec %s: %s\n", argv[i], strerror(errno));
      _exit(127);
    }
    pids.push_back(pid);
  }
  int status_out = 0;
  size_t left = pids.size();
  while (left > 0) {
    int st = 0;
    const pid_t done = wait(&st);
    if (done < 0) break;
    --left;
    const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
    if (code != 0 && status_out == 0) {
      status_out = code;
      for (pid_t p : pids)
        if (p != done) kill(p, SIGTERM);  // exact pids created above
    }
  }
  const std::string cleanup = std::string("rm -rf ") + dir;
  if (system(cleanup.c_str()) != 0) fprintf(stderr, "spnrun: could not remove %s\n", dir);
  return status_out;
}

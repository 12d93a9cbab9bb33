// `ba` and `slam` -- drop-in command-line tools for the reference's two host
// programs (ba/ba.cpp:479-1085, ba/slam.cpp:479-1135), re-hosted on the C ABI of
// include/gbp_cuda.h + include/gbp_host.h.  One source, two binaries
// (-DGBP_CLI_SLAM selects the incremental-SLAM schedule).
//
// Kept from the reference: every flag with its default and help text
// (ba.cpp:400-465, slam.cpp:400-465), the input format, the schedule, and the
// log lines a user greps.  Changed: `--ipus N` is the
// number of GPUs (one process per GPU: rank 0 spawns the others and broadcasts
// the NCCL id through the environment), `--camspertile` is accepted and ignored,
// the per-sweep metric comes from the device (24 bytes per sweep instead of a
// full belief read-back), `--profile true` writes gbp_profile.json.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/gbp_cuda.h"
#include "../../include/gbp_host.h"

#ifdef GBP_CLI_SLAM
static const bool kSlam = true;
#else
static const bool kSlam = false;
#endif

namespace {

struct Flag {
  const char* name;
  const char* arg;  // "arg" or "arg (=default)"
  const char* help;
};

const Flag kFlags[] = {
    {"help", "", "Show command help"},
    {"bal_file", "arg", "Set the bal file"},
#ifdef GBP_CLI_SLAM
    {"iters_between_kfs", "arg (=700)", "Number of iterations of synchronous GBP between adding successive keyframes"},
#else
    {"n_iters", "arg (=1500)", "Number of iterations of synchronous GBP"},
#endif
    {"profile", "arg (=0)", "Save profile report after execution"},
    {"ipus", "arg (=1)", "Number of GPUs to use (one process per GPU; the reference's number of IPU chips)"},
    {"camspertile", "arg (=1)", "Accepted for compatibility with the IPU tile mapping; ignored"},
    {"tn", "arg (=0)", "Set keyframe translation noise value"},
    {"rn", "arg (=0)", "Set keyframe rotation noise value"},
    {"ltn", "arg (=0)", "Set landmark translation noise noise value"},
    {"avdepth_on", "arg (=0)",
     "bool: should landmarks be initialised at an average depth from the keyframe they are first observed by"},
    {"avdepth", "arg (=1)",
     "float: Average depth at which landmarks are initialed from the keyframe which they are first observed by."},
    {"reproj_meas_var", "arg (=4)",
     "Variance of Gaussian noise in Gaussian measurement model for the reprojection constraints"},
    {"prior_std_weaker_factor", "arg (=100)",
     "Factor: std of gauss noise of reprojection factors / std of gauss noise of prior factors"},
    {"first_cam_prior_std", "arg (=0.00999999978)",
     "Standard deviation of prior on pose of first keyframe, to anchor optimisation."},
    {"steps", "arg (=5)", "The priors are gradually weakened over this many steps."},
    {"undamped_start", "arg (=15)", "Number of undamped iterations before damping GBP."},
    {"v", "arg (=0)", "Verbose: print beliefs"},
    {"seed", "arg (=0)", "Seed of the --tn/--rn/--ltn noise (0 = clock, like the reference)"},
#ifndef GBP_CLI_SLAM
    {"converge", "arg (=0)",
     "If > 0: after the prior weakening, stop as soon as the reprojection error improves by less than this (relative) "
     "over 10 sweeps, or exceeds twice its running minimum; --n_iters is then the maximum"},
#else
    {"host_keyframes", "arg (=0)",
     "bool: insert keyframes through the reference's READ_PRIORS / host / NEW_KEYFRAME round trip instead of on the "
     "device (same result bit for bit)"},
#endif
    {"out", "arg", "Write the optimised problem (belief means) to this file, in the input format"},
};

void print_help() {
  std::cout << "Options:\n";
  for (const Flag& f : kFlags) {
    std::string left = std::string("  --") + f.name + (f.arg[0] ? " " : "") + f.arg;
    if (left.size() < 38) left.resize(38, ' ');
    else left += "\n" + std::string(38, ' ');
    std::cout << left << f.help << "\n";
  }
  std::cout << "\n";
}

struct ParseError {
  std::string msg;
};

bool parse_bool(const std::string& name, const std::string& v) {
  std::string s;
  for (char c : v) s += (char)std::tolower((unsigned char)c);
  if (s == "1" || s == "true" || s == "yes" || s == "on") return true;
  if (s == "0" || s == "false" || s == "no" || s == "off") return false;
  throw ParseError{"the argument ('" + v + "') for option '--" + name + "' is invalid. Valid choices are 'on|off', 'yes|no', '1|0' and 'true|false'"};
}
template <class T>
T parse_num(const std::string& name, const std::string& v) {
  char* end = nullptr;
  const double d = std::strtod(v.c_str(), &end);
  if (v.empty() || *end) throw ParseError{"the argument ('" + v + "') for option '--" + name + "' is invalid"};
  return (T)d;
}

struct Cli {
  gbp_cli_options opt;
  std::string bal_file, out_file;
  float converge = 0.f;
  bool host_keyframes = false;
  bool help = false;
};

Cli parse(int argc, char** argv) {
  Cli c;
  gbp_cli_options_default(&c.opt);
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) throw ParseError{"too many positional options have been specified on the command line"};
    a = a.substr(2);
    std::string val;
    bool has_val = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) {
      val = a.substr(eq + 1);
      a = a.substr(0, eq);
      has_val = true;
    }
    const Flag* flag = nullptr;
    for (const Flag& f : kFlags)
      if (a == f.name) flag = &f;
    if (!flag) throw ParseError{"unrecognised option '--" + a + "'"};
    if (a == "help") {
      c.help = true;
      continue;
    }
    if (!has_val) {
      if (i + 1 >= argc) throw ParseError{"the required argument for option '--" + a + "' is missing"};
      val = argv[++i];
    }
    gbp_cli_options& o = c.opt;
    if (a == "bal_file") c.bal_file = val;
    else if (a == "n_iters") o.n_iters = parse_num<int>(a, val);
    else if (a == "iters_between_kfs") o.iters_between_kfs = parse_num<int>(a, val);
    else if (a == "profile") o.profile = parse_bool(a, val);
    else if (a == "ipus") o.n_ipus = parse_num<int>(a, val);
    else if (a == "camspertile") o.cams_per_tile = parse_num<int>(a, val);
    else if (a == "tn") o.transnoise = parse_num<float>(a, val);
    else if (a == "rn") o.rotnoise = parse_num<float>(a, val);
    else if (a == "ltn") o.lmktrans_noise = parse_num<float>(a, val);
    else if (a == "avdepth_on") o.av_depth_on = parse_bool(a, val);
    else if (a == "avdepth") o.av_depth = parse_num<float>(a, val);
    else if (a == "reproj_meas_var") o.reproj_meas_var = parse_num<float>(a, val);
    else if (a == "prior_std_weaker_factor") o.prior_std_weaker_factor = parse_num<float>(a, val);
    else if (a == "first_cam_prior_std") o.first_cam_prior_std = parse_num<float>(a, val);
    else if (a == "steps") o.steps = parse_num<float>(a, val);
    else if (a == "undamped_start") o.iters_before_damping = parse_num<int>(a, val);
    else if (a == "v") o.verbose = parse_bool(a, val);
    else if (a == "seed") o.noise_seed = parse_num<uint32_t>(a, val);
    else if (a == "converge") c.converge = parse_num<float>(a, val);
    else if (a == "host_keyframes") c.host_keyframes = parse_bool(a, val);
    else if (a == "out") c.out_file = val;
  }
  return c;
}

// ---- multi-GPU launch: rank 0 spawns ranks 1..N-1 of the same command line ---------
struct Ranks {
  int world = 1, rank = 0;
  unsigned char nccl_id[128];
  std::vector<pid_t> children;
};

std::string to_hex(const unsigned char* p, size_t n) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (size_t i = 0; i < n; ++i) {
    s += d[p[i] >> 4];
    s += d[p[i] & 15];
  }
  return s;
}

bool setup_ranks(Ranks& r, int n_gpus, char** argv) {
  const char* er = std::getenv("GBP_CLI_RANK");
  const char* ew = std::getenv("GBP_CLI_WORLD");
  const char* ei = std::getenv("GBP_CLI_NCCL_ID");
  if (er && ew && ei && std::strlen(ei) == 256) {  // a spawned rank
    r.rank = std::atoi(er);
    r.world = std::atoi(ew);
    for (int i = 0; i < 128; ++i) {
      unsigned v = 0;
      std::sscanf(ei + 2 * i, "%2x", &v);
      r.nccl_id[i] = (unsigned char)v;
    }
    return true;
  }
  r.world = n_gpus;
  if (n_gpus <= 1) return true;
  if (gbp_cuda_nccl_unique_id(r.nccl_id) != GBP_OK) {
    std::cerr << "ERROR: " << gbp_cuda_last_error() << "\n";
    return false;
  }
  setenv("GBP_CLI_WORLD", std::to_string(n_gpus).c_str(), 1);
  setenv("GBP_CLI_NCCL_ID", to_hex(r.nccl_id, 128).c_str(), 1);
  for (int k = 1; k < n_gpus; ++k) {
    const pid_t pid = fork();
    if (pid == 0) {
      setenv("GBP_CLI_RANK", std::to_string(k).c_str(), 1);
      execv("/proc/self/exe", argv);
      std::perror("execv");
      _exit(127);
    }
    r.children.push_back(pid);
  }
  return true;
}

struct Out {  // only rank 0 talks
  bool on;
  template <class T>
  Out& operator<<(const T& v) {
    if (on) std::cout << v;
    return *this;
  }
};

void print_beliefs(gbp_handle* h, uint32_t C, uint32_t L) {  // --v true  (ba.cpp:1030-1051)
  std::vector<float> ce(6 * (size_t)C), cl(36 * (size_t)C), le(3 * (size_t)L), ll(9 * (size_t)L);
  gbp_cuda_get_beliefs(h, ce.data(), cl.data(), le.data(), ll.data(), nullptr, nullptr, nullptr);
  std::cout << "\nKeyframe Eta beliefs: \n";
  for (unsigned i = 0; i < 6 && 6 + i < ce.size(); ++i) std::printf("%.12f  ", ce[6 + i]);
  std::cout << "\nKeyframe Lambda beliefs: \n";
  for (unsigned i = 0; i < 36 && 36 + i < cl.size(); ++i) std::printf("%.12f  ", cl[36 + i]);
  std::cout << '\n';
  std::cout << "\nLandmark Eta beliefs: \n";
  for (unsigned i = 0; i < 12 && i < le.size(); ++i) std::printf("%.12f  ", le[i]);
  std::cout << "\nLandmark Lambda beliefs: \n";
  for (unsigned i = 0; i < 18 && i < ll.size(); ++i) std::printf("%.12f  ", ll[i]);
  std::cout << '\n';
  std::fflush(stdout);
}

#define CHECK(expr)                                                     \
  do {                                                                  \
    if ((expr) != GBP_OK) {                                             \
      std::cerr << "ERROR: " << #expr << ": " << gbp_cuda_last_error() << "\n"; \
      return 2;                                                         \
    }                                                                   \
  } while (0)

int run(const Cli& cli, Ranks& rk) {
  const gbp_cli_options& options = cli.opt;
  Out out{rk.rank == 0};
  gbp_bal* bal = nullptr;
  if (gbp_bal_load(cli.bal_file.c_str(), &bal) != GBP_OK) {
    std::cerr << "ERROR: unable to open file " << cli.bal_file << "\n";  // ba.cpp:484-487
    return 1;
  }
  gbp_setup* setup = nullptr;
  CHECK(gbp_setup_create(bal, &options, kSlam ? GBP_MODE_SLAM : GBP_MODE_BA, &setup));
  const gbp_problem* p = gbp_setup_problem(setup);
  const uint32_t n_keyframes = p->n_keyframes, n_points = p->n_points, n_edges = p->n_edges;
  out << (kSlam ? "Loaded data onto host!\n" : "Completed loading data!\n");
  out << (kSlam ? "SLAM\n" : "\nBundle Adjustment\n");
  out << "\nNumber of keyframe nodes in the graph: " << n_keyframes << '\n';
  out << "Number of landmark nodes in the graph: " << n_points << '\n';
  out << "Number of edges in the graph: " << n_edges << '\n';
  out << "\nNumber of GPUs: " << rk.world << '\n';
  if (kSlam && rk.world > 1) {
    std::cerr << "ERROR: incremental SLAM runs on one GPU (--ipus 1)\n";
    return 1;
  }

  out << "\nAttaching to GPU device...\n";
  gbp_opts o;
  gbp_opts_default(&o);
  o.device = rk.rank;
  gbp_handle* h = nullptr;
  const auto time0 = std::chrono::steady_clock::now();
  double kf_seconds = 0.0;
  unsigned kf_inserted = 0;
  out << "Running program to stream initial data to GPU\n";
  const int rc = (rk.world > 1) ? gbp_cuda_init_shard(p, &o, (uint32_t)rk.world, (uint32_t)rk.rank, rk.nccl_id, &h)
                                : gbp_cuda_init(p, &o, &h);
  if (rc == GBP_ERR_CUDA) {
    std::cout << "Could not find a device\n" << gbp_cuda_last_error() << "\n";  // ba.cpp:652-655
    std::exit(-1);
  }
  if (rc != GBP_OK) {
    std::cerr << "ERROR: " << gbp_cuda_last_error() << "\n";
    return 2;
  }
  out << "Attached to device: " << o.device << "\n";
  out << "Initial data streaming complete\n\n";
  out << "Sending priors and computing factor potentials.\n";
  uint32_t Cl = 0, Ll = 0, El = 0, mk = 0, ml = 0;
  gbp_cuda_dims(h, &Cl, &Ll, &El, &mk, &ml);
  if (options.profile) gbp_cuda_set_profile(h, 1);

  gbp_iter_stats s;
  CHECK(gbp_cuda_eval(h, &s));
  out << "Initial Reprojection error: " << s.reproj_mean << " Cost " << s.cost << "\n";

  const unsigned steps2 = (unsigned)(options.steps * 2);
  double device_ms = 0, factor_ms = 0, variable_ms = 0;
  uint64_t sweeps = 0, launches = 0;
  auto account = [&](int n) {
    float ms = 0, a = 0, b = 0;
    uint64_t k = 0;
    gbp_cuda_last_timing(h, &ms, &k);
    gbp_cuda_last_kernel_times(h, &a, &b);
    device_ms += ms; factor_ms += a; variable_ms += b; launches += k; sweeps += (uint64_t)n;
  };
  std::vector<gbp_iter_stats> st;
  // Sweeps between two schedule events are enqueued together; the per-sweep metric is
  // evaluated on the device and printed afterwards, line for line as the reference does.
  auto sweep_block = [&](unsigned iter0, unsigned n, unsigned iter_base) -> int {
    const unsigned chunk = options.verbose ? 1u : 64u;
    for (unsigned done = 0; done < n;) {
      const unsigned m = std::min(chunk, n - done);
      st.resize(m);
      CHECK(gbp_cuda_iterate(h, (int)m, st.data()));
      account((int)m);
      for (unsigned k = 0; k < m; ++k) {
        const unsigned iter = iter0 + done + k;
        if (kSlam)
          out << "Iters " << iter_base + iter << " (since last kf " << iter << ") // Reprojection error " << st[k].reproj_mean;
        else
          out << "Iter " << iter << " // Reprojection error " << st[k].reproj_mean;
        out << " // Cost " << st[k].cost << " // n relins: " << st[k].n_relins << " // n robust edges " << st[k].n_robust << "\n";
        if (options.verbose && rk.rank == 0) print_beliefs(h, Cl, Ll);
      }
      done += m;
    }
    return 0;
  };
  // one stretch of the schedule: `n` sweeps starting at local iteration `iter`, with the
  // prior weakening of ba.cpp:1003-1006 / slam.cpp:1049-1052 in front of sweeps 1,3,5,...
  auto stretch = [&](unsigned iter, unsigned n, unsigned iter_base) -> int {
    const unsigned end = iter + n;
    while (iter < end) {
      if ((iter + 1) % 2 == 0 && iter < steps2) {
        out << "Weakening priors \n";
        CHECK(gbp_cuda_weaken_priors(h));
      }
      const unsigned m = (iter >= steps2) ? end - iter : 1u;  // past the weakening phase: one block
      if (int e = sweep_block(iter, m, iter_base)) return e;
      iter += m;
    }
    return 0;
  };

  if (!kSlam) {
    out << "Number of iterations: " << options.n_iters << "\n";
    const unsigned n_total = (unsigned)std::max(options.n_iters, 0);
    if (cli.converge > 0.f && n_total > steps2) {
      if (int e = stretch(0, steps2, 0)) return e;  // the weakening phase runs as scheduled
      st.resize(n_total - steps2);
      int n_done = 0, why = 0;
      CHECK(gbp_cuda_iterate_until(h, (int)(n_total - steps2), 10, cli.converge, 2.0f, st.data(), &n_done, &why));
      account(n_done);
      for (int k = 0; k < n_done; ++k)
        out << "Iter " << steps2 + k << " // Reprojection error " << st[k].reproj_mean << " // Cost " << st[k].cost
            << " // n relins: " << st[k].n_relins << " // n robust edges " << st[k].n_robust << "\n";
      out << "Stopped after " << steps2 + n_done << " iterations: "
          << (why == GBP_STOP_CONVERGED ? "converged" : why == GBP_STOP_DIVERGED ? "diverging" : "iteration limit") << "\n";
    } else if (int e = stretch(0, n_total, 0)) {
      return e;
    }
  } else {
    const unsigned ibk = (unsigned)std::max(options.iters_between_kfs, 1);
    const unsigned niters = (n_keyframes - 1) * ibk - 1;  // slam.cpp:1013
    out << "Total number of GBP iterations: " << niters << "\n";
    out << "GBP iterations between sucessive keyframes: " << ibk << "\n";
    std::vector<float> cbe(6 * (size_t)n_keyframes), cbl(36 * (size_t)n_keyframes), cpe(cbe.size()), cpl(cbl.size()),
        lpe(3 * (size_t)n_points), lpl(9 * (size_t)n_points);
    std::vector<int32_t> dcount(n_edges);
    unsigned i = 0, data_counter = 0;
    while (i < niters) {
      if ((i + 1) % ibk == 0) {  // slam.cpp:1020-1046
        int n_new = 0;
        const auto t_kf = std::chrono::steady_clock::now();
        if (cli.host_keyframes) {
          CHECK(gbp_cuda_get_beliefs(h, cbe.data(), cbl.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
          CHECK(gbp_cuda_get_priors(h, cpe.data(), cpl.data(), lpe.data(), lpl.data()));  // READ_PRIORS
          CHECK(gbp_setup_next_keyframe(setup, cbe.data(), cbl.data(), cpe.data(), cpl.data(), lpe.data(), lpl.data(),
                                        dcount.data(), &n_new));
          data_counter = (unsigned)gbp_setup_data_counter(setup);
          const gbp_problem* q = gbp_setup_problem(setup);
          CHECK(gbp_cuda_add_keyframe(h, dcount.data(), cpe.data(), cpl.data(), lpe.data(), lpl.data(), q->active_flag,
                                      q->cam_weaken_flag, q->lmk_weaken_flag));  // NEW_KEYFRAME
        } else {  // the same insertion where the data lives: 4 bytes cross the bus
          data_counter += 1;
          CHECK(gbp_cuda_add_keyframe_device(h, data_counter + 1, (uint32_t)cli.opt.steps, &n_new));
        }
        kf_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_kf).count();
        kf_inserted += 1;
        out << "\n**********************************************************";
        out << "\n Adding keyframe " << data_counter + 1;
        out << "\n Adding " << n_new << " new landmarks";
        out << "\n**********************************************************\n\n";
      }
      // sweeps up to (not including) the next insertion point
      const unsigned next = ((i + 1) % ibk == 0) ? i + ibk : (i / ibk + 1) * ibk - 1;
      const unsigned n = std::min(niters, next) - i;
      if (int e = stretch(0, n, ibk * data_counter)) return e;
      i += n;
    }
  }
  out << "\n Finished GBP.\n";
  if (!cli.out_file.empty() && rk.world == 1) {
    std::vector<float> ce(6 * (size_t)n_keyframes), cl(36 * (size_t)n_keyframes), le(3 * (size_t)n_points), ll(9 * (size_t)n_points);
    CHECK(gbp_cuda_get_beliefs(h, ce.data(), cl.data(), le.data(), ll.data(), nullptr, nullptr, nullptr));
    gbp_bal* opt = nullptr;
    CHECK(gbp_bal_with_means(bal, ce.data(), cl.data(), le.data(), ll.data(), &opt));
    const int wrc = gbp_bal_save(opt, cli.out_file.c_str());
    gbp_bal_free(opt);
    if (wrc != GBP_OK) {
      std::cerr << "ERROR: " << gbp_cuda_last_error() << "\n";
      return 2;
    }
    out << "Optimised problem written to " << cli.out_file << "\n";
  }
  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - time0).count();
  if (rk.rank == 0) {
    std::printf("Timing report: wall %.3f s; %llu sweeps, device time in sweeps %.3f ms (%.2f us/sweep, %.1f sweeps/s, "
                "%.4g factor-message updates/s), %llu kernel launches\n",
                wall, (unsigned long long)sweeps, device_ms, sweeps ? 1e3 * device_ms / sweeps : 0.0,
                device_ms > 0 ? 1e3 * sweeps / device_ms : 0.0, device_ms > 0 ? 1e3 * (double)n_edges * sweeps / device_ms : 0.0,
                (unsigned long long)launches);
    if (kf_inserted)
      std::printf("Keyframe insertions: %u, %.3f ms each (%s)\n", kf_inserted, 1e3 * kf_seconds / kf_inserted,
                  cli.host_keyframes ? "READ_PRIORS / host / NEW_KEYFRAME round trip" : "on the device");
    if (options.profile) {
      const char* log_dir = std::getenv("GC_PROFILE_LOG_DIR");  // same variable as the reference (ba.cpp:1062)
      const std::string path = std::string(log_dir ? log_dir : ".") + "/gbp_profile.json";
      std::ofstream f(path);
      f << "{\"sweeps\": " << sweeps << ", \"device_ms\": " << device_ms << ", \"factor_kernel_ms\": " << factor_ms
        << ", \"variable_kernel_ms\": " << variable_ms << ", \"kernel_launches\": " << launches << ", \"wall_s\": " << wall
        << ", \"n_gpus\": " << rk.world << ", \"factors\": " << n_edges << "}\n";
      std::cout << "Profile written to " << path << "\n";
    }
  }
  gbp_cuda_free(h);
  gbp_setup_free(setup);
  gbp_bal_free(bal);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Cli cli;
  try {
    cli = parse(argc, argv);
  } catch (const ParseError& e) {
    std::cerr << "error: " << e.msg << "\n";
    return 1;
  }
  if (cli.help) {  // the reference prints the options and then throws (ba.cpp:469-472)
    print_help();
    return 1;
  }
  if (cli.bal_file.empty()) {
    std::cerr << "error: the option '--bal_file' is required but missing\n";
    return 1;
  }
  Ranks rk;
  int n_gpus = cli.opt.n_ipus <= 0 ? 1 : cli.opt.n_ipus;
  if (!setup_ranks(rk, n_gpus, argv)) return 2;
  int rc = run(cli, rk);
  for (pid_t pid : rk.children) {
    int status = 0;
    waitpid(pid, &status, 0);
    if (rc == 0 && (!WIFEXITED(status) || WEXITSTATUS(status) != 0)) rc = 3;
  }
  return rc;
}

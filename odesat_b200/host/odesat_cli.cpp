// odesat_cli.cpp — `solve` / `batch` / `inter` / `stoch` driver over libodesat_b200 with the reference's
// flags (-f -o -t -n -s -l -b, main.rs:31-141) and console lines (main.rs:156-200, 263-320,
// 335-383).  SURVEY §8f row 1: the DIMACS reader, normaliser and result writer are restated here
// (cnf.rs:138-219, 246-264, 289-315) so the GPU path is runnable end to end without the Rust crate.
// `solve` runs the reference's ratio preprocessing first (`-r`, default 7.0, main.rs:150-166; restated in
// preprocess.hpp) and replays the elimination trace on the result (main.rs:186-187).  `stoch` (main.rs:206-252) runs
// the same preprocessing and then the weighted random-flip search of src/stoch.rs on the GPU (-b replicas; default 1).
// Extra flags: --seed (the reference's RNG is OS-seeded), --f32, --gpus N (replica shards of `batch` /
// `inter` over N devices of this process; default: every visible device), --chunk K (steps between early-exit polls).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>

#include "preprocess.hpp"
#include "system.hpp"

using namespace odesat;

// cnf.rs:138-172
static CNFFormula parse_dimacs_format(const std::string& input, std::vector<std::vector<int>>& raw) {
    CNFFormula f;
    bool have = false;
    std::istringstream in(input);
    std::string line;
    std::set<std::size_t> vars;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == 'c') continue;
        if (line.rfind("p cnf", 0) == 0) {
            std::istringstream ls(line);
            std::string p, cnf;
            ls >> p >> cnf >> f.varnum;
            have = true;
            continue;
        }
        std::istringstream ls(line);
        std::string tok;
        std::vector<int> lits;
        while (ls >> tok) {
            if (tok == "0") break;
            lits.push_back(std::stoi(tok));
        }
        raw.push_back(lits);
        for (int l : lits) vars.insert((std::size_t)std::abs(l));
    }
    if (!have) f.varnum = vars.size();
    return f;
}

// cnf.rs:206-219 (ascending file name instead of HashSet order; varnum stays the header value)
static std::map<std::size_t, std::size_t> normalize(const std::vector<std::vector<int>>& raw, CNFFormula& f) {
    std::set<std::size_t> vars;
    for (const auto& c : raw) for (int l : c) vars.insert((std::size_t)std::abs(l));
    std::map<std::size_t, std::size_t> name_map;
    for (std::size_t v : vars) { const std::size_t k = name_map.size(); name_map[v] = k; }
    for (const auto& c : raw) {
        CNFClause cl;
        for (int l : c) cl.literals.push_back({name_map[(std::size_t)std::abs(l)], l < 0});
        f.clauses.push_back(cl);
    }
    return name_map;
}

// cnf.rs:246-264 on the ORIGINAL formula; missing variables are inserted as false (`entry().or_insert(false)`)
static bool evaluate_cnf(std::map<std::size_t, bool>& values, const std::vector<std::vector<int>>& raw) {
    for (const auto& c : raw) {
        bool sat = false;
        for (int l : c) {
            const bool val = values.emplace((std::size_t)std::abs(l), false).first->second;
            sat = sat || (l < 0 ? !val : val);
        }
        if (!sat) return false;
    }
    return true;
}

static int usage() {
    std::fprintf(stderr, "usage: odesat_b200_cli <solve|stoch|batch|inter> -f FILE [-o OUT] [-t TOL] [-n STEPS] [-s STEP] [-l ZETA] [-r RATIO] [-b BATCH] [--seed S] [--f32] [--gpus N] [--chunk K]\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    const std::string cmd = argv[1];
    std::string input, output;
    std::optional<double> tol, step, zeta;
    std::optional<std::size_t> steps;
    std::size_t batch = 0;
    uint64_t seed = 1;
    bool f32 = false;
    int gpus = 0, chunk = 0;                                                // 0: every visible device / library default
    double ratio = 7.0;                                                     // main.rs:150-154
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "-f" || a == "--input") input = next();
        else if (a == "-o" || a == "--output") output = next();
        else if (a == "-t" || a == "--tolerance") tol = std::atof(next());
        else if (a == "-n" || a == "--step-number") steps = (std::size_t)std::atoll(next());
        else if (a == "-s" || a == "--step-size") step = std::atof(next());
        else if (a == "-l" || a == "--learning-rate") zeta = std::atof(next());
        else if (a == "-b" || a == "--batch-size") batch = (std::size_t)std::atoll(next());
        else if (a == "--seed") seed = (uint64_t)std::atoll(next());
        else if (a == "--f32") f32 = true;
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--chunk") chunk = std::atoi(next());
        else if (a == "-r" || a == "--ctv-ratio") ratio = std::atof(next());
        else return usage();
    }
    if (cmd == "preprocess" && !input.empty()) {
        // host-only helper (no GPU): print the reduced formula of `solve`'s preprocessing in DIMACS form and the
        // elimination trace, one step per line — used by the CPU tests to compare with odesat_b200/preprocess.py
        std::ifstream fh(input);
        if (!fh) { std::perror(input.c_str()); return 1; }
        std::stringstream ss;
        ss << fh.rdbuf();
        std::vector<std::vector<int>> raw;
        CNFFormula f = parse_dimacs_format(ss.str(), raw);
        prep::ClauseSet set;
        for (const auto& c : raw) set.insert(prep::make_clause(c));
        const prep::Trace trace = prep::repeatedly_resolve_and_update(set, f.varnum, (float)ratio);
        auto print_clause = [](const prep::Clause& c) {
            for (prep::Lit l : c) std::printf("%s%zu ", prep::neg_of(l) ? "-" : "", prep::var_of(l));
            std::printf("0\n");
        };
        std::printf("p cnf %zu %zu\n", f.varnum, set.size());
        for (const auto& c : set) print_clause(c);
        for (const auto& st : trace) {
            std::printf("t %s %zu %zu\n", st.ve ? "ve" : "bce", st.var, st.clauses.size());
            for (const auto& c : st.clauses) { std::printf("t  "); print_clause(c); }
        }
        return 0;
    }
    if (input.empty() || (cmd != "solve" && cmd != "batch" && cmd != "inter" && cmd != "stoch")) return usage();
    if ((cmd == "batch" || cmd == "inter") && batch == 0) return usage();
    if (cmd == "batch" && !steps) return usage();                       // main.rs:96-97: -n is required
    try {
        std::printf("Reading CNF formula from file...\n");
        std::ifstream fh(input);
        if (!fh) { std::perror(input.c_str()); return 1; }
        std::stringstream ss;
        ss << fh.rdbuf();
        std::printf("Parsing CNF formula...\n");
        std::vector<std::vector<int>> raw;
        CNFFormula f = parse_dimacs_format(ss.str(), raw);
        prep::Trace trace;
        std::vector<std::vector<int>> integrated = raw;                     // the clauses the ODE is built from
        if (cmd == "solve" || cmd == "stoch") {                             // main.rs:162-166, 223-226
            std::printf("Preprocessing CNF formula...\n");
            prep::ClauseSet set;
            for (const auto& c : raw) set.insert(prep::make_clause(c));
            trace = prep::repeatedly_resolve_and_update(set, f.varnum, (float)ratio);
            integrated.clear();                                             // convert_to_cnf_formula: BTreeSet order
            for (const auto& c : set) {
                std::vector<int> lits;
                for (prep::Lit l : c) lits.push_back(prep::neg_of(l) ? -(int)prep::var_of(l) : (int)prep::var_of(l));
                integrated.push_back(lits);
            }
        } else {
            std::printf("Normalizing CNF formula...\n");
        }
        const auto name_map = normalize(integrated, f);
        system::Formula F(f);
        std::printf("Simulating...\n");
        odesat_params p = system::make_params(tol, step, steps, zeta);
        p.precision = f32 ? ODESAT_F32 : ODESAT_F64;
        p.chunk = chunk;
        const int64_t R = cmd == "solve" ? 1 : (cmd == "stoch" ? (int64_t)std::max<std::size_t>(batch, 1) : (int64_t)batch);
        // replicas are independent (main.rs:278-308, system.rs:279-289): shard them over the devices of this process
        p.n_gpus = cmd == "solve" ? 1 : (gpus > 0 ? gpus : std::max(1, odesat_device_count()));
        const int mode = cmd == "inter" ? ODESAT_MODE_INTER : ODESAT_MODE_BATCH;
        if (cmd == "solve" && !steps) p.steps = 1 << 30;   // the reference loops forever on UNSAT input
        std::vector<uint8_t> assignment(f.varnum), verified((std::size_t)R);
        std::vector<int64_t> solved((std::size_t)R);
        int64_t winner = -1, run = 0;
        // states are generated on the device (main.rs:283-289 with a seeded generator)
        const auto t0 = std::chrono::steady_clock::now();
        if (cmd == "stoch")   // stoch.rs:80-110 search, R independent replicas from the reference's initial state
            system::check(odesat_stoch_search(F.handle(), R, nullptr, nullptr, seed, 0, steps ? (int64_t)*steps : -1, chunk, 0,
                                              solved.data(), verified.data(), &winner, assignment.data(), &run));
        else
            system::check(odesat_simulate_batch(F.handle(), R, nullptr, nullptr, nullptr, seed, 0, &p, mode, 0, solved.data(),
                                                verified.data(), &winner, assignment.data(), &run));
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::map<std::size_t, bool> values;                              // cnf.rs:301-315
        for (const auto& kv : name_map) values[kv.first] = assignment[kv.second] != 0;
        if (cmd == "solve" || cmd == "stoch") {
            std::printf("Mapping values...\n");
            prep::calculate_trace(values, trace);                           // main.rs:186-187
            std::printf("Evaluating CNF formula...\n");
        }
        const bool ok = evaluate_cnf(values, raw);
        std::printf("%sChecking if solution vector satisfies formula: %s\n", (cmd == "solve" || cmd == "stoch") ? "" : "\n", ok ? "true" : "false");
        std::fprintf(stderr, "[odesat_b200] replicas=%lld steps_run=%lld winner=%lld\n", (long long)R, (long long)run, (long long)winner);
        // one JSON record of the integration (the driver reports throughput next to the reference's lines)
        std::fprintf(stderr, "{\"variables\": %zu, \"clauses\": %zu, \"replicas\": %lld, \"steps_run\": %lld, \"seconds\": %.6f, "
                             "\"clause_evals_per_s\": %.4g, \"precision\": \"%s\", \"gpus\": %d}\n",
                     f.varnum, f.clauses.size(), (long long)R, (long long)run, sec,
                     sec > 0 ? (double)run * (double)f.clauses.size() * (double)R / sec : 0.0, f32 ? "f32" : "f64",
                     (int)std::min<int64_t>(p.n_gpus, R));
        std::printf("Rendering variable assignments...\n");
        std::string render;                                              // cnf.rs:289-298
        for (const auto& kv : values) render += std::to_string(kv.first) + " " + (kv.second ? "1" : "0") + "\n";
        if (!output.empty()) {
            std::printf("Writing results to file...\n");
            std::ofstream(output) << render;
        } else {
            std::printf("Variable assignments:\n%s\n", render.c_str());
        }
        return 0;
    } catch (const system::Error& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}

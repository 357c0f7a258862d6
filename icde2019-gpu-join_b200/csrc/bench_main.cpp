// bench_main.cpp -- command-line benchmark driver of the B200 radix-join engine.
//
// Same command line and stdout lines as the reference driver src/main.cu (flags :445-457,
// "INPUT:" echo :443-554, relation-creation messages :186-262, algorithm line :267), so runs of
// the two binaries diff cleanly.  Work is done by the reference-shaped operator entry point
// hashJoinClusteredProbe (include/gpujoin_operator.h) through the algorithm table, as
// main.cu:64,291 does.  Extra flags (not in the reference): --seed-r/--seed-s (the reference
// seeds from time(NULL)), --payload ones|rowid, --parallel-gen, --repeat.
#include <getopt.h>

#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gpujoin.h"
#include "gpujoin_generator.h"
#include "gpujoin_operator.h"

namespace {

struct Algorithm {
    const char* name;
    unsigned int (*run)(args*, timingInfo*);
};
const Algorithm kAlgorithms[] = {{"HJC", hashJoinClusteredProbe}};

struct Options {
    int option = 0;
    const Algorithm* alg = nullptr;
    uint64_t nR = 0, nS = 0;
    double skew = 0.0;
    bool unique = true, full_range = false, from_files = false, parallel_gen = false;
    std::string fileR, fileS;
    int threads = 32, values_per_thread = 2, shared_mem = 30720, pivots = 1, one_to_many = 0;
    uint64_t multS = 1, multR = 1;
    unsigned int seedR = 1, seedS = 2;
    std::string payload = "ones";
    int repeat = 1;
};

[[noreturn]] void usage(int code) {
    fprintf(stderr,
            "usage: bench -b 7 -a HJC -R <tuples> -S <tuples> [-s <zipf>] [--non-unique] [--full-range]\n"
            "             [--file -k <R.bin> -l <S.bin>] [-x <S copies>] [-y <R copies>]\n"
            "             [--seed-r N] [--seed-s N] [--payload ones|rowid] [--parallel-gen] [--repeat N]\n"
            "  -b 7 runs the join, -b 8 only generates the relations\n");
    exit(code);
}

Options parse(int argc, char** argv) {
    Options o;
    int non_unique = 0, full_range = 0, file_in = 0, par = 0;
    const option longopts[] = {
        {"non-unique", no_argument, &non_unique, 1}, {"full-range", no_argument, &full_range, 1},
        {"file", no_argument, &file_in, 1},          {"parallel-gen", no_argument, &par, 1},
        {"benchmark", required_argument, nullptr, 'b'}, {"alg", required_argument, nullptr, 'a'},
        {"SelsNum", required_argument, nullptr, 'S'},   {"RelsNum", required_argument, nullptr, 'R'},
        {"skew", required_argument, nullptr, 's'},      {"threadsNum", required_argument, nullptr, 't'},
        {"valuesPerThread", required_argument, nullptr, 'v'}, {"sharedMem", required_argument, nullptr, 'm'},
        {"pivotsNum", required_argument, nullptr, 'p'}, {"OneToMany", no_argument, nullptr, 'w'},
        {"seed-r", required_argument, nullptr, 1001},   {"seed-s", required_argument, nullptr, 1002},
        {"payload", required_argument, nullptr, 1003},  {"repeat", required_argument, nullptr, 1004},
        {nullptr, 0, nullptr, 0}};
    printf("INPUT: ");
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "b:a:k:l:S:R:s:t:v:m:p:x:y:", longopts, &idx)) != -1) {
        switch (c) {
            case 0: printf("%s\t", longopts[idx].name); break;
            case 'b': o.option = atoi(optarg); printf("option = %d\t", o.option); break;
            case 'a':
                for (const Algorithm& a : kAlgorithms)
                    if (!strcmp(a.name, optarg)) o.alg = &a;
                if (!o.alg) { fprintf(stderr, "unknown algorithm %s\n", optarg); usage(1); }
                printf("joinAlg = %s\t", o.alg->name);
                break;
            case 'k': o.fileR = optarg; printf("R filename = %s\t", optarg); break;
            case 'l': o.fileS = optarg; printf("S filename = %s\t", optarg); break;
            case 'S': o.nS = strtoull(optarg, nullptr, 10); printf("||S|| = %lu\t", (unsigned long)o.nS); break;
            case 'R': o.nR = strtoull(optarg, nullptr, 10); printf("||R|| = %lu\t", (unsigned long)o.nR); break;
            case 's': o.skew = atof(optarg); printf("skew = %f\t", o.skew); break;
            case 't': o.threads = atoi(optarg); printf("#threads = %d\t", o.threads); break;
            case 'v': o.values_per_thread = atoi(optarg); printf("values per thread= %d\t", o.values_per_thread); break;
            case 'm': o.shared_mem = atoi(optarg); printf("sharedMem = %d\t", o.shared_mem); break;
            case 'p': o.pivots = atoi(optarg); printf("pivotsNum = %d\t", o.pivots); break;
            case 'w': o.one_to_many = 1; printf("OneToMany = %d\t", o.one_to_many); break;
            case 'x': o.multS = strtoull(optarg, nullptr, 10); printf("SelsMultiplier = %lu\t", (unsigned long)o.multS); break;
            case 'y': o.multR = strtoull(optarg, nullptr, 10); printf("RelsMultiplier = %lu\t", (unsigned long)o.multR); break;
            case 1001: o.seedR = (unsigned)strtoul(optarg, nullptr, 10); break;
            case 1002: o.seedS = (unsigned)strtoul(optarg, nullptr, 10); break;
            case 1003: o.payload = optarg; break;
            case 1004: o.repeat = atoi(optarg); break;
            default: usage(1);
        }
    }
    printf("\n");
    o.unique = !non_unique; o.full_range = full_range; o.from_files = file_in; o.parallel_gen = par;
    if (o.option != 7 && o.option != 8) usage(0);
    if (o.option == 7 && !o.alg) { fprintf(stderr, "-a HJC is required\n"); usage(1); }
    if (o.from_files && (o.fileR.empty() || o.fileS.empty())) { fprintf(stderr, "--file needs -k and -l\n"); usage(1); }
    return o;
}

std::string cache_name(const char* fmt, uint64_t a, uint64_t b = 0) {
    char buf[96];
    snprintf(buf, sizeof(buf), fmt, (unsigned long)a, (unsigned long)b);
    return buf;
}

}  // namespace

int main(int argc, char** argv) {
    Options o = parse(argc, argv);
    const uint64_t baseR = o.nR, baseS = o.nS;
    const uint64_t nR = o.nR * o.multR, nS = o.nS * o.multS;
    const uint64_t mbR = nR * sizeof(int) / 1024 / 1024, mbS = nS * sizeof(int) / 1024 / 1024;

    int *R = nullptr, *S = nullptr;
    if (gj_malloc_pinned((void**)&R, nR * sizeof(int)) || gj_malloc_pinned((void**)&S, nS * sizeof(int))) {
        fprintf(stderr, "Problem allocating space for the relations: %s\n", gj_last_error());
        return 1;
    }
    std::vector<int> baseBufR, baseBufS;
    int* genR = R; int* genS = S;
    if (o.multR > 1) { baseBufR.resize(baseR); genR = baseBufR.data(); }
    if (o.multS > 1) { baseBufS.resize(baseS); genS = baseBufS.data(); }

    if (o.from_files) {
        printf("Reading from files\n");
        if (gj_read_relation(o.fileR.c_str(), R, nR) || gj_read_relation(o.fileS.c_str(), S, nS)) {
            fprintf(stderr, "\ncould not read the relation files\n");
            return 1;
        }
    } else if (o.full_range) {
        printf("Creating relation R with %lu tuples (%lu MB) using non-unique keys and full range : ", (unsigned long)nR, (unsigned long)mbR);
        fflush(stdout);
        gj_seed_generator(o.seedR);
        gj_create_relation_nonunique(cache_name("pk_R%lu.bin", nR).c_str(), R, nR, INT_MAX);
        printf("Creating relation S with %lu tuples (%lu MB) using non-unique keys and full range : ", (unsigned long)nS, (unsigned long)mbS);
        fflush(stdout);
        gj_create_relation_fk_from_pk(cache_name("fk_S%lu_pk_R%lu.bin", nS, nR).c_str(), S, nS, R, nR);
    } else if (o.unique) {
        printf("Creating relation R with %lu tuples (%lu MB) using unique keys : ", (unsigned long)nR, (unsigned long)mbR);
        fflush(stdout);
        // unlike main.cu:135,143 the cache names carry the seed, so R and S never alias one file
        if (o.parallel_gen) gj_create_relation_unique_parallel(genR, baseR, (int64_t)baseR, o.seedR, 0);
        else gj_create_relation_unique(cache_name("unique_%lu_seed%lu.bin", baseR, o.seedR).c_str(), genR, baseR, (int64_t)baseR, o.seedR);
        if (o.multR > 1) gj_create_relation_n(genR, R, baseR, o.multR);
        if (o.skew > 0) {
            printf("Creating relation S with %lu tuples (%lu MB) using unique keys and skew %f : ", (unsigned long)nS, (unsigned long)mbS, o.skew);
            fflush(stdout);
            const uint64_t alphabet = o.multS > 1 ? baseS : nR;
            if (o.parallel_gen) gj_create_relation_zipf_parallel(genS, baseS, (unsigned)alphabet, o.skew, o.seedS, 0);
            else {
                gj_seed_generator(o.seedS);
                char nm[96];
                snprintf(nm, sizeof(nm), "unique_skew%.2f_S%lu_seed%u.bin", o.skew, (unsigned long)baseS, o.seedS);
                gj_create_relation_zipf(nm, genS, baseS, (int64_t)alphabet, o.skew);
            }
        } else {
            printf("Creating relation S with %lu tuples (%lu MB) using unique keys : ", (unsigned long)nS, (unsigned long)mbS);
            fflush(stdout);
            const int64_t maxid = (int64_t)(o.multS > 1 ? baseS : nR);
            if (o.parallel_gen) gj_create_relation_unique_parallel(genS, baseS, maxid, o.seedS, 0);
            else gj_create_relation_unique(cache_name("unique_S%lu_max%lu", baseS, (uint64_t)maxid).append("_seed").append(std::to_string(o.seedS)).append(".bin").c_str(), genS, baseS, maxid, o.seedS);
        }
        if (o.multS > 1) gj_create_relation_n(genS, S, baseS, o.multS);
    } else {
        printf("Creating relation R with %lu tuples (%lu MB) using non-unique keys : ", (unsigned long)nR, (unsigned long)mbR);
        fflush(stdout);
        gj_seed_generator(o.seedR);
        gj_create_relation_nonunique(cache_name("nonUnique_R%lu.bin", nR).c_str(), R, nR, (int64_t)(nR / 2));
        printf("Creating relation S with %lu tuples (%lu MB) using non-unique keys : ", (unsigned long)nS, (unsigned long)mbS);
        fflush(stdout);
        gj_create_relation_nonunique(cache_name("nonUnique_S%lu.bin", nS).c_str(), S, nS, (int64_t)(nR / 2));
    }
    fflush(stdout);

    int rc = 0;
    if (o.option == 7) {
        printf("%s : shareMemory = %ld\t#threads = %d\n", o.alg->name, (long)o.shared_mem, o.threads);
        fflush(stdout);
        for (int it = 0; it < o.repeat; ++it) {
            unsigned int r;
            if (o.payload == "rowid") {
                std::vector<int> Pr(nR), Ps(nS);
                for (uint64_t i = 0; i < nR; ++i) Pr[i] = (int)i;
                for (uint64_t i = 0; i < nS; ++i) Ps[i] = (int)i;
                r = outOfGPU_Join1_payload(R, Pr.data(), nR, S, Ps.data(), nS, nullptr, 0, 0, 0);
            } else {
                args a;
                memset(&a, 0, sizeof(a));
                a.R = R; a.R_els = nR; a.S = S; a.S_els = nS;
                a.threadsNum = o.threads; a.sharedMem = (unsigned)o.shared_mem; a.pivotsNum = (unsigned)o.pivots;
                r = o.alg->run(&a, nullptr);
            }
            if (r == ~0u) { rc = 2; break; }
            const gj_operator_result* res = gj_operator_last_result();
            printf("matches %lu checksum %lu\n", (unsigned long)res->matches, (unsigned long)res->checksum);
        }
        gj_operator_release();
    }
    gj_free_pinned(R);
    gj_free_pinned(S);
    return rc;
}

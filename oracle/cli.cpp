// ORACLE CLI (test infrastructure): accepts the subset of `pandora map` argv that
// /root/reference/src/lib.rs:594-617 + src/predict.rs:288-294 build, so it can stand in for the
// pandora executable via drprg's -p/--pandora option (src/predict.rs:137-144).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
extern "C" {
int orc_run_map(const char*, const char*, const char*, const char*, const void*, uint32_t, uint32_t);
const char* orc_last_error();
}
struct Opts {
    uint32_t threads = 1, min_cluster_size = 10;
    uint8_t illumina = 0, debug = 0;
    uint32_t genome_size = 5000000, max_covg = 300;
    double gt_conf = 1, genotyping_error_rate = 0.01;
    uint32_t max_diff = 0;
    double error_rate = 0;
};
int main(int argc, char** argv) {
    if (argc < 2 || strcmp(argv[1], "map") != 0) {
        fprintf(stderr, "usage: pandora_oracle map [pandora map options] <prg> <reads>\n");
        return 2;
    }
    Opts o;
    uint32_t w = 14, k = 15;
    std::string outdir = "pandora", vcf_refs;
    std::vector<std::string> pos;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "--genotype" || a == "--local" || a == "-v" || a == "-vv") continue;
        else if (a == "--gt-conf" || a == "-G") o.gt_conf = atof(val());
        else if (a == "-o" || a == "--outdir") outdir = val();
        else if (a == "-g" || a == "--genome-size") o.genome_size = (uint32_t)strtoul(val(), 0, 10);
        else if (a == "--max-covg") o.max_covg = (uint32_t)strtoul(val(), 0, 10);
        else if (a == "--vcf-refs") vcf_refs = val();
        else if (a == "-t" || a == "--threads") o.threads = (uint32_t)atoi(val());
        else if (a == "-w") w = (uint32_t)atoi(val());
        else if (a == "-k") k = (uint32_t)atoi(val());
        else if (a == "-c" || a == "--min-cluster-size") o.min_cluster_size = (uint32_t)atoi(val());
        else if (a == "-m" || a == "--max-diff") o.max_diff = (uint32_t)atoi(val());
        else if (a == "-e" || a == "--error-rate") o.error_rate = atof(val());
        else if (a == "-E" || a == "--gt-error-rate") o.genotyping_error_rate = atof(val());
        else if (a == "-I" || a == "--illumina") o.illumina = 1;
        else if (a == "-K" || a == "--debugging-files") o.debug = 1;
        else if (a[0] == '-') { fprintf(stderr, "pandora_oracle: unsupported option %s\n", a.c_str()); return 2; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { fprintf(stderr, "pandora_oracle: need <prg> <reads>\n"); return 2; }
    std::string cmd = "mkdir -p '" + outdir + "'";
    if (system(cmd.c_str()) != 0) return 1;
    int rc = orc_run_map(pos[0].c_str(), pos[1].c_str(), vcf_refs.c_str(), outdir.c_str(), &o, w, k);
    if (rc) fprintf(stderr, "pandora_oracle: %s\n", orc_last_error());
    return rc;
}

// sim_geom.cpp -- the reference's CLI (src/main.cpp) on the B200 engine: same flags, same params.conf,
// same .geom files, same phase timers on stdout.
#include <chrono>
#include <iostream>

#include "sj_host.hpp"

int main(int argc, char **argv) {
    parse_settings args;
    int ret = parse_args(&args, &argc, argv);
    if (ret) return ret;
    int precision = SJ_F64, n_sets = 2, n_gpus = 1;
    for (int i = 1; i < argc; ++i) {                       // engine-only extras, not in the reference
        if (argv[i] && !strcmp(argv[i], "--fp32")) precision = SJ_F32;
        if (argv[i] && !strcmp(argv[i], "--real-fields")) n_sets = 1;
        if (argv[i] && !strcmp(argv[i], "--gpus") && i + 1 < argc && argv[i + 1]) n_gpus = atoi(argv[i + 1]);
    }
    char *name = args.conf_fname ? args.conf_fname : strdup("params.conf");
    ret = parse_conf_file(&args, name);
    if (!args.conf_fname) free(name);
    if (ret) return ret;
    correct_defaults(&args);

    parse_ercode ercode = E_SUCCESS;
    auto start = std::chrono::steady_clock::now();
    sj_bound_geom geom(args, &ercode, precision, n_sets, 1, n_gpus);
    if (ercode) return (int)ercode;
    auto end_init = std::chrono::steady_clock::now();
    std::cout << "initialization completed in " << std::chrono::duration_cast<std::chrono::milliseconds>(end_init - start).count()
              << " ms (GPU rasterization " << geom.raster_ms << " ms)" << std::endl;
    geom.run(args.out_dir);
    auto end_run = std::chrono::steady_clock::now();
    std::cout << "simulation completed in " << std::chrono::duration<double>(end_run - end_init).count() << " s ("
              << geom.n_t_pts << " steps)" << std::endl;
    geom.save_field_times(args.out_dir);
    auto end_write = std::chrono::steady_clock::now();
    std::cout << "saving timeseries completed in " << std::chrono::duration_cast<std::chrono::milliseconds>(end_write - end_run).count() << " ms" << std::endl;
    std::cout << "total time: " << std::chrono::duration<double>(end_write - start).count() << " s" << std::endl;
    cleanup_settings(&args);
    return 0;
}

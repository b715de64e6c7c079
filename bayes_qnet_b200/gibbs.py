#!/usr/bin/env python3
"""Front end for running the sampler: the reference's ``gibbs.py`` (gibbs.py:1-248) with the same options and the
same output files, on the GPU engine.  New options: ``--chains`` (independent chains; files are written for chain 0,
R-hat / ESS over all chains are printed), ``--device``, ``--sampler SLICE|MH``.

    python -m bayes_qnet_b200.gibbs --network net.yml --arrivals arrv.yml --bayes 1 --gibbs-iter 1000 --chains 1024
"""
import getopt
import sys
import time

import numpy

from . import estimation, qnet, qnetu, qstats, sampling

USAGE = """Usage:
     gibbs.py <options>
    Front-end for running the queueing network Gibbs sampler.
     --network (-n)       YAML file describing network (required)
     --arrivals (-a)      YAML file describing arrivals (required)
     --multiarr (-A)      Three-file prefix describing arrivals (alternative to -a)
     --arrvtbl            Arrivals file in table format
     --burn <int>         Number of iterations to burn in (default 100)
     --bayes <T|F>        If true, use Bayes instead of StEM
     --departures-only    If supplied, hold params constant; only sample departure times
     --gibbs-iter <int>   Number of iterations to sample (default 100)
     --output-mu <file>   If supplied, write series of parameter vectors to file
     --output-arrv <prefix> If supplied, write arrivals to files with given prefix
     --arrv-iter <int>    Number of iterations at which to output arrivals
     --thin <int>         Number of sweeps per iteration in stochastic EM
     --statistics <list>  List of statistics to report after every iteration.
                          Example: --stat "mean_wait, mean_qsize_arriving"
     --stats-iter <int>   After estimation, sample this many more iterations and report the statistics
     --do-init <int>      Whether to run initialization (True by default)
     --params FILE        Read in parameters from file
     --config FILE        YAML configuration (sampler, initializer, seed: qnetu.read_configuration)
     --seed <int>         Random seed
     --chains <int>       Independent chains run on the GPU (default 1)
     --device <int>       CUDA device (default 0)
     --sampler <name>     SLICE (default) or MH
         Available statistics: %s
"""


def usage():
    print(USAGE % qstats.all_stats_string())


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    nburn = niter = 100
    netf = arrf = multif = arrvtbl = None
    output_mu = output_arrv = config_file = params = None
    arrv_iter = 100
    all_stats = []
    stats_iter = 0
    thin = 1
    do_bayes = 0
    do_init = 1
    departures_only = False
    sweeten = 0
    chains, device = 1, 0
    try:
        opts, args = getopt.getopt(argv, "hn:a:A:", [
            "help", "seed=", "network=", "arrivals=", "gibbs-iter=", "burn=", "output-mu=", "output-arrv=", "config=",
            "statistics=", "stats=", "thin=", "stats-iter=", "multiarr=", "arrvtbl=", "bayes=", "departures-only",
            "arrv-iter=", "do-init=", "sweeten=", "params=", "chains=", "device=", "sampler="])
    except getopt.GetoptError as err:
        print(str(err))
        usage()
        return 2
    for o, a in opts:
        if o in ("-h", "--help"):
            usage()
            return 0
        elif o in ("-n", "--network"): netf = a
        elif o in ("-a", "--arrivals"): arrf = a
        elif o in ("-A", "--multiarr"): multif = a
        elif o == "--arrvtbl": arrvtbl = a
        elif o == "--do-init": do_init = int(a)
        elif o == "--burn": nburn = int(a)
        elif o == "--sweeten": sweeten = int(a)
        elif o == "--gibbs-iter": niter = int(a)
        elif o == "--output-mu": output_mu = a
        elif o == "--output-arrv": output_arrv = a
        elif o == "--arrv-iter": arrv_iter = int(a)
        elif o == "--thin": thin = int(a)
        elif o == "--seed": sampling.set_seed(int(a))
        elif o == "--config": config_file = a
        elif o == "--departures-only": departures_only = True
        elif o in ("--statistics", "--stats"):
            if a == "ALL":
                all_stats = qstats.STATS[:]
            else:
                all_stats = [getattr(qstats, name.strip()) for name in a.split(",") if name.strip()]
        elif o == "--stats-iter": stats_iter = int(a)
        elif o == "--bayes": do_bayes = int(a)
        elif o == "--params": params = a
        elif o == "--chains": chains = int(a)
        elif o == "--device": device = int(a)
        elif o == "--sampler": estimation.set_sampler(a)

    t0 = time.time()
    if config_file:
        with open(config_file) as f:
            qnetu.read_configuration(f)
    if netf is None:
        print("ERROR: Specify --network")
        return 1
    net = qnetu.qnet_from_fname(netf)
    if params:
        with open(params) as f:
            net.parameters = [float(x) for x in f.readlines()[0].split()]
    if arrf is not None:
        with open(arrf) as f:
            arrv = qnetu.load_arrivals(net, f)
    elif multif is not None:
        arrv = qnetu.read_multif_of_prefix(multif, net)
    elif arrvtbl is not None:
        arrv = qnetu.read_from_table(net, arrvtbl)
    else:
        raise Exception("Must specify either -a or -A")

    print(net.as_yaml())
    print("Number of events = ", arrv.num_events())
    print("Number hidden events = ", arrv.num_hidden_events())
    print("Thin = ", thin)
    if do_init:
        arrv = net.gibbs_initialize(arrv)
    if output_arrv:
        with open("initial.txt", "w") as f:
            arrv.write_csv(f)
            f.write("\n")
    arrv.validate()

    mu_f = open(output_mu, "w") if output_mu else None
    train_stats_f = [open("train_%s.txt" % fn.__name__, "w") for fn in all_stats]

    def output_train_stats(net_, arrv_, it, lp=None):
        if output_arrv and it % arrv_iter == 0:
            with open("%s%d.txt" % (output_arrv, it), "w") as f:
                arrv_.write_csv(f)
                f.write("\n")
        for f, fn in zip(train_stats_f, all_stats):
            f.write("%d %s\n" % (it, qstats.stringify(fn(arrv_))))
            f.flush()

    # per-iteration files need the chain-0 state on the host every iteration; without them the whole run is one launch
    report = output_train_stats if (output_arrv or all_stats) else None
    kw = dict(n_chains=chains, device=device)
    if do_bayes:
        mul, arrvl = estimation.bayes(net, arrv, niter, report_fn=report, sweeten=sweeten, **kw)
    elif departures_only:
        mul, arrvl = estimation.sample_departures(net, arrv, niter, report_fn=report, **kw)
    else:
        mul, arrvl = estimation.sem(net, arrv, nburn, niter, gibbs_iter=thin, report_fn=report, **kw)
    if mu_f:
        for mu in mul:
            mu_f.write(" ".join(map(str, mu)) + "\n")
        mu_f.close()
    for f in train_stats_f:
        f.close()

    run = estimation.last_run()
    if run is not None:
        st = run.stats
        print("Events resampled per sec = ", st["nevents"] / max(run.seconds, 1e-9))
        if chains > 1 and run.theta.shape[0] >= 4:
            print("RHAT ", " ".join("%.4f" % x for x in run.rhat()))
            print("ESS ", " ".join("%.1f" % x for x in run.ess()))

    if stats_iter > 0:  # statistics of a fresh sample at the final parameters (gibbs.py:219-236)
        test_stats_f = [open("%s.txt" % fn.__name__, "w") for fn in all_stats]

        def output_test_stats(net_, arrv_, it):
            for f, fn in zip(test_stats_f, all_stats):
                f.write(qstats.stringify(fn(arrv_)) + "\n")

        net.parameters = mul[-1]
        net.gibbs_resample(arrvl[-1], 0, stats_iter, return_arrv=False, report_fn=output_test_stats)
        for f in test_stats_f:
            f.close()

    lp = net.log_prob(arrvl[-1])
    n = numpy.sum(qstats.nobs_by_queue(arrvl[-1]))
    print("LOG_PROB ", lp)
    print("AIC ", 2 * len(net.parameters) - 2 * lp)
    print("BIC ", numpy.log(max(n, 1)) * len(net.parameters) - 2 * lp)
    print("Total time ", time.time() - t0)
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""YAML model loading, arrivals file formats and configuration -- the reference's ``qnetu`` surface
(qnetu.py:16-569) restated for Python 3 over the structure-of-arrays ``Arrivals``.

Model files are unchanged: ``states:`` (name, queues, successors) and ``queues:`` (name, service, optional
``type``/``processors``), e.g. ``examples/models/mm1.yml`` in the reference.
"""
import numpy
import yaml

from . import arrivals, distributions, estimation, hmm, qnet, queues, sampling


def qnet_from_fname(fname):
    with open(fname) as f:
        return qnet_from_text(f)


def qnet_from_text(yaml_text):
    """Build a Qnet from a YAML description (qnetu.py:22-91)."""
    ylist = yaml.safe_load(yaml_text)
    if not isinstance(ylist, dict):
        raise Exception("Bad YAML file.  Needed a dictionary:\n%s" % ylist)
    qname2id, sname2id = {}, {}
    snames, qnames = [], []
    s_succ, s_q = {}, {}

    def add_state(name):
        if name not in sname2id:
            snames.append(name)
            sname2id[name] = len(sname2id)

    def add_queue(name):
        if name not in qname2id:
            qnames.append(name)
            qname2id[name] = len(qname2id)

    for s in ylist["states"]:
        add_state(s["name"])
        sid = sname2id[s["name"]]
        if "successors" in s:
            for x in s["successors"]:
                add_state(x)
            s_succ[sid] = [sname2id[x] for x in s["successors"]]
        for x in s["queues"]:
            add_queue(x)
        s_q[sid] = [qname2id[x] for x in s["queues"]]

    ns, nq = len(snames), len(qnames)
    T = numpy.zeros((ns, ns))
    O = numpy.zeros((ns, nq))
    for sid, succl in s_succ.items():
        for succ in succl:
            T[sid, succ] = 1. / len(succl)  # uniform routing (qnetu.py:63-67)
    for sid, ql in s_q.items():
        for qid in ql:
            O[sid, qid] = 1. / len(ql)
    fsm = hmm.HMM(T, O)

    qs, factors, universe = [], [], {}
    by_name = {}
    for qhash in ylist["queues"]:
        q = queue_from_yhash(qhash, universe)
        by_name[q.name] = q
    for name in qnames:  # queue ids follow first mention in `states`, like the reference's qname2id
        if name not in by_name:
            raise Exception("Queue %s is used by a state but not described under queues:" % name)
    # the reference indexes self.queues by position in the `queues:` list and qname2id by first mention in
    # `states:`; its own models list them in the same order, which we require
    listed = [qh["name"] for qh in ylist["queues"]]
    if listed != qnames:
        raise Exception("queues: must list the queues in the order the states mention them: %s vs %s" % (listed, qnames))
    for name in qnames:
        qs.append(by_name[name])
        factors.append(by_name[name].service)
    return qnet.Qnet(qs, factors, universe, qname2id, sname2id, fsm)


def queue_from_text(text, universe):
    return queue_from_yhash(yaml.safe_load(text), universe)


def queue_from_yhash(qhash, universe):
    """qnetu.py:96-122: GG1 -> QueueGGk(k=1), GGK, DELAY, PS, GG1R."""
    qname = qhash["name"]
    sdist = _tmpl_from_yaml(qhash["service"], qname, universe)
    num_procs = qhash.get("processors", 1)
    if "type" in qhash:
        qtype = qhash["type"]
    elif num_procs == 1:
        qtype = "GG1"
    else:
        qtype = "GGk"
    qtype = qtype.upper()
    if qtype == "GG1":
        return queues.QueueGGk(qname, sdist, 1)
    if qtype == "GGK":
        return queues.QueueGGk(qname, sdist, num_procs)
    if qtype == "DELAY":
        return queues.DelayStation(qname, sdist)
    if qtype == "PS":
        return queues.QueuePS(qname, sdist)
    if qtype == "GG1R":
        return queues.QueueGG1R(qname, sdist)
    raise Exception("Could not understand queue %s type %s" % (qname, qtype))


def _parse_args(tag):
    if isinstance(tag, list):
        return tag[0], tag[1:]
    return tag, None


def _tmpl_from_yaml(tag, qname, universe):
    dist_type, args = _parse_args(tag)
    if dist_type == "MIX":  # [MIX, w0, [M, 3.0], w1, [LN, 0.0, 1.0], ...]   (qnetu.py:142-155)
        ws, dists = [], []
        for i in range(0, len(args), 2):
            ws.append(args[i])
            subtype, subargs = _parse_args(args[i + 1])
            dists.append(_dist_from_yaml(subtype, subargs))
        var_name = "COMPONENT_%s" % qname
        if universe is not None:
            universe[var_name] = list(range(len(ws)))
        return queues.MixtureTemplate(var_name, ws, dists)
    return queues.DistributionTemplate(_dist_from_yaml(dist_type, args))


def _dist_from_yaml(dist_type, args):
    """qnetu.py:160-186."""
    if dist_type == "M":
        return distributions.Exponential(args[0] if args else 1.0)
    if dist_type == "G":
        if isinstance(args, list) and len(args) == 1 and isinstance(args[0], dict):
            args = args[0]
        if isinstance(args, list):
            shape, scale = args
        elif isinstance(args, dict):
            shape, scale = args["shape"], args["scale"]
        else:
            shape = scale = 1.0
        return distributions.Gamma(shape=shape, scale=scale)
    if dist_type == "LN":
        if isinstance(args, list) and len(args) == 1 and isinstance(args[0], dict):
            args = args[0]
        if isinstance(args, list):
            mu, sig = args
        elif isinstance(args, dict):
            mu, sig = args["mu"], args["sigma"]
        else:
            mu, sig = 0, 1
        return distributions.LogNormal(mu, sig)
    if dist_type == "D":
        return distributions.Dirac(args[0])
    raise Exception("Could not understand tag %s" % dist_type)


# ---- arrivals formats -------------------------------------------------------------------------------
def _arrivals_from_dicts(net, evt_dict_lst):
    """!Arrivals constructor (qnetu.py:244-305): sort by (a, d), append, recompute services, drop non-FIFO tasks."""
    rows = []
    for eid, l in enumerate(evt_dict_lst):
        q = net.queue_by_name(l["queue"])
        if q is None:
            raise Exception("Error in YAML file: Could not find queue $%s$" % l["queue"])
        st = l.get("state", -1)
        sid = st if isinstance(st, int) else net.sid_by_name(st)
        obs = l.get("obs", 1)
        rows.append((float(l["arrival"]), float(l["departure"]), int(l["task"]), net.qid_of_queue(q), sid,
                     int(l.get("obs_a", obs)), int(l.get("obs_d", obs)), eid))
    rows.sort(key=lambda r: (r[0], r[1]))
    if not rows:
        return arrivals.Arrivals(net)
    a, d, tid, qid, sid, oa, od, eid = map(list, zip(*rows))
    arrv = arrivals.Arrivals.from_arrays(net, tid, qid, sid, a, d, oa, od, eid=eid)
    # Remove non-FIFO tasks (qnetu.py:291-303)
    while True:
        bad = numpy.nonzero(arrv.soa()["s"] < 0)[0]
        if len(bad) == 0:
            break
        e = arrivals.Event._view(arrv, bad[0])
        prev = e.previous_by_queue()
        victim = prev.tid if prev is not None else e.tid
        print("WARNING: Deleting non-fifo task %d" % victim)
        arrv.delete_task(victim)
    arrv.validate()
    return arrv


class _ArrivalsLoader(yaml.SafeLoader):
    pass


_ArrivalsLoader.add_constructor(u"!Event", lambda loader, node: loader.construct_mapping(node))
_ArrivalsLoader.add_constructor(u"!Arrivals", lambda loader, node: loader.construct_mapping(node, deep=True))


def load_arrivals(net, stream):
    """Read the ``!Arrivals {events: [!Event {...}]}`` format (qnetu.py:317, examples/models/arrv.yml)."""
    m = yaml.load(stream, Loader=_ArrivalsLoader)
    if m is None:
        return None
    return _arrivals_from_dicts(net, m["events"])


def arrv_from_yaml(net, yaml_text):
    """Plain list of event dictionaries (qnetu.py:189-227)."""
    return _arrivals_from_dicts(net, yaml.safe_load(yaml_text))


def write_multifile_arrv(arrv, statef, af, df, obs_only=True):
    """qnetu.py:343-356: three files -- structure, arrivals, departures -- 15 decimals."""
    for evt in arrv:
        if (not obs_only) or (evt.obs_a or evt.obs_d):
            nxt = evt.next_by_task()
            statef.write("%d %d %s %d %d\n" % (evt.tid, evt.eid, evt.queue().name, evt.state, nxt.eid if nxt else -1))
    for evt in arrv:
        if (not obs_only) or evt.obs_a:
            af.write("%d %.15f %d\n" % (evt.eid, evt.a, evt.obs_a))
    for evt in arrv:
        if (not obs_only) or evt.obs_d:
            df.write("%d %.15f %d\n" % (evt.eid, evt.d, evt.obs_d))


def read_multifile_arrv(net, statef, af, df):
    """qnetu.py:358-452."""
    e2a, e2d = {}, {}
    for f, dst in ((af, e2a), (df, e2d)):
        for line in f.readlines():
            fields = line.split()
            if not fields:
                continue
            if len(fields) == 2:
                fields.append(1)
            dst[int(fields[0])] = (float(fields[1]), int(fields[2]))
    evts = []
    nxt_of = {}
    for line in statef.readlines():
        if not line.split():
            continue
        tid, eid, qname, sid, nxt = line.split()
        eid, tid, sid, nxt = int(eid), int(tid), int(sid), int(nxt)
        q = net.queue_by_name(qname)
        if q is None:
            raise Exception("Cannot find queue named %s in\n%s" % (qname, net.as_yaml()))
        if eid not in e2a:
            raise Exception("No arrival file entry for event %d" % eid)
        if eid not in e2d:
            raise Exception("No arrival file entry for event %d" % eid)
        a, obs_a = e2a[eid]
        d, obs_d = e2d[eid]
        evts.append((d, a, net.qid_of_queue(q), tid, eid, sid, obs_a, obs_d))
        nxt_of[eid] = nxt
    evts.sort()
    # by-task order: follow the next-event links of the state file
    by_eid = dict((e[4], e) for e in evts)
    has_prev = set(n for n in nxt_of.values() if n >= 0)
    ordered = []
    for e in evts:
        if e[4] in has_prev:
            continue
        cur = e[4]
        while cur is not None and cur >= 0 and cur in by_eid:
            ordered.append(by_eid[cur])
            cur = nxt_of.get(cur, -1)
    d, a, qid, tid, eid, sid, oa, od = map(list, zip(*ordered)) if ordered else ([],) * 8
    return arrivals.Arrivals.from_arrays(net, tid, qid, sid, a, d, oa, od, eid=eid)


def write_multif_to_prefix(prefix, arrv, obs_only=True):
    with open("%sstate.txt" % prefix, "w") as sf, open("%sa.txt" % prefix, "w") as af, open("%sd.txt" % prefix, "w") as df:
        write_multifile_arrv(arrv, sf, af, df, obs_only=obs_only)


def read_multif_of_prefix(prefix, net):
    with open("%sstate.txt" % prefix) as sf, open("%sa.txt" % prefix) as af, open("%sd.txt" % prefix) as df:
        return read_multifile_arrv(net, sf, af, df)


def read_csv_from_lines(net, lines):
    """The table ``Arrivals.write_csv`` produces (qnetu.py:476-506)."""
    rows = []
    for line in lines[1:]:
        f = line.split()
        if not f:
            continue
        rows.append((int(f[0]), int(f[1]), int(f[2]), float(f[3]), float(f[4]), int(f[7]), int(f[8]), int(f[9]), int(f[10])))
    rows.sort(key=lambda r: r[3])
    if not rows:
        return arrivals.Arrivals(net)
    eid, tid, qid, a, d, state, proc, oa, od = map(list, zip(*rows))
    return arrivals.Arrivals.from_arrays(net, tid, qid, state, a, d, oa, od, eid=eid, proc=proc)


def read_from_table(net, fname):
    with open(fname) as f:
        return read_csv_from_lines(net, f.readlines())


def read_table_from_string(net, string):
    return read_csv_from_lines(net, string.split("\n"))


def write_to_table(fname, arrv):
    with open(fname, "w") as f:
        arrv.write_csv(f)


# ---- configuration (qnetu.py:527-562) ---------------------------------------------------------------------
def read_configuration(stream):
    yml = yaml.safe_load(stream)
    if not isinstance(yml, dict):
        raise Exception("Could not understand YAML in config file.")
    for k in yml.keys():
        if k == "one_d_sampler":
            qnet.set_sampler(yml[k])
        elif k == "debug_integrate":
            pass
        elif k == "sampling_type":
            estimation.set_sampler(yml[k])
        elif k == "use_rtp":
            qnet.set_use_rtp(int(yml[k]))
        elif k == "proposal":
            if str(yml[k]).lower() != "twos":
                raise NotImplementedError("MH proposal %s (only the default, twos, is built)" % yml[k])
        elif k == "initializer":
            qnet.set_initialization_type(yml[k])
        elif k == "slice_sorted":
            qnet.set_slice_sorted(int(yml[k]))
        else:
            raise Exception("Could not understand config key %s" % k)


def set_seed(seed):
    """qnetu.py:567."""
    sampling.set_seed(seed)

#!/usr/bin/env python3
"""phyloligo.py on B200: same command line, file formats and dispatcher functions
as the reference script (reference phylopackage/bin/phyloligo.py), with the scoop /
joblib workers replaced by the CUDA library.

Drop-in seam:
    compute_frequencies(...) -> (frequencies, freq_name)      reference :980-997
    compute_distances(...)   -> res | None                    reference :536-553
and the worker-level functions compute_frequency (:663), frequency_pack (:795),
compute_frequency_memmap (:693), compute_frequency_h5py_chunk (:756),
compute_unpack (:166), distances_loc (:195), distances_h5py (:233).

``--method`` and ``-c`` are accepted for compatibility; they only select the
return conventions of the reference's back-ends, never a CPU code path.

Deliberate fixes of reference defects: SC works (the
reference raises NameError); KT/BC/SC work in every output mode; memmap mode
computes the matrix once; -q works with --large h5py.
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

from . import engine, io_formats, phylodist
from ._lib import PhyloligoError, METRICS, TILE

PANEL_ROWS = None  # rows per streamed panel; None = engine.auto_panel_rows (about 256 MB per pinned buffer)


def remove_folder(folder):
    try:
        shutil.rmtree(folder)
    except Exception:
        print("Failed to delete folder: {}".format(folder))


# ---------------------------------------------------------------------------
# one process per GPU (torchrun): the reference's n_jobs workers become the ranks of one node
# ---------------------------------------------------------------------------
def ranks():
    """(rank, world) of this process.  Under torchrun (WORLD_SIZE > 1) the NCCL process group is
    created on first use, one GPU per rank; a plain ``python phyloligo.py`` run is (0, 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch.distributed as dist
    if not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        engine.bind_host_to_gpu_node(local)  # pinned panels in the memory next to this GPU
    return dist.get_rank(), dist.get_world_size()


def _barrier():
    if ranks()[1] > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()


def _broadcast_str(value):
    """rank 0's string on every rank (temp-dir names must agree)."""
    if ranks()[1] == 1:
        return value
    import torch.distributed as dist
    box = [value]
    dist.broadcast_object_list(box, src=0)
    return box[0]


# ---------------------------------------------------------------------------
# worker-level API (what scoop / joblib call in the reference)
# ---------------------------------------------------------------------------
def compute_frequency(seq, pattern="1111", strand="both"):
    """Frequency vector of one sequence (reference :663-691)."""
    _check_strand(strand)
    res = engine.profile_sequences([seq], str(pattern), strand, want=("freq64",))
    return res["freq64"][0].cpu().numpy()


def frequency_pack(params):
    """reference :795-813"""
    return compute_frequency(*params)


def compute_frequency_memmap(frequency, i, seq, pattern="1111", strand="both"):
    """Store row i of a caller-owned (memmap) array (reference :693-720)."""
    frequency[i] = compute_frequency(seq, pattern, strand)


def compute_frequency_h5py_chunk(freq_name_folder, seqchunk, pattern, strand, start, stop):
    """One chunk of sequences -> file frequencies_{start}_{stop}, dataset
    'frequencies', float32 (reference :756-792)."""
    _check_strand(strand)
    res = engine.profile_sequences(list(seqchunk), str(pattern), strand, want=("freq32",))
    freqs = res["freq32"].cpu().numpy()
    path = os.path.join(freq_name_folder, "frequencies_{}_{}".format(start, stop))
    io_formats.write_hdf5(path, "frequencies", freqs)


def compute_unpack(params):
    """(i, j, freqi, freqj, distname) -> (i, j, d)   (reference :166-171)"""
    i, j, freqi, freqj, distname = params
    _check_metric(distname)
    return i, j, engine.pair_distance(freqi, freqj, distname)


def _large_metric(metric):
    """The reference's --large workers compute Eucl through sklearn's Gram-form
    euclidean_distances (:200-202, :238-246); here that is the tensor-core kernel."""
    return "EuclGram" if metric == "Eucl" else metric


def distances_loc(output, X, s, metric):
    """output[s] = D(X[s], X) for one block-row slice (reference :195-222)."""
    _check_metric(metric)
    metric = _large_metric(metric)
    device = engine.require_cuda()
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(device)
    P, aux, dim = engine.prepare(Xd, metric)
    n = int(P.shape[0])
    start, stop, _ = s.indices(n)
    out_dtype = torch.float32 if np.dtype(output.dtype) == np.float32 else torch.float64
    blk = torch.empty((stop - start, n), dtype=out_dtype, device=device)
    engine.distance_block(metric, P, aux, dim, start, stop, 0, n, blk, start, 0)
    output[s] = blk.cpu().numpy()


def distances_h5py(output_dir, input_path, s, metric):
    """Block row of the matrix from an HDF5 frequency file into
    distance_{start}_{stop}, dataset 'distances', float32 (reference :233-301)."""
    _check_metric(metric)
    X = io_formats.read_hdf5(input_path, "frequencies")
    n = X.shape[0]
    start, stop, _ = s.indices(n)
    out = np.empty((stop - start, n), dtype=np.float32)
    holder = _RowHolder(out, start)
    distances_loc(holder, X, slice(start, stop), metric)
    path = os.path.join(output_dir, "distance_{}_{}".format(start, stop))
    io_formats.write_hdf5(path, "distances", out)


class _RowHolder:
    """Adapter so distances_loc can fill a block-local array through output[s]."""

    def __init__(self, arr, row0):
        self.arr, self.row0, self.dtype = arr, row0, arr.dtype

    def __setitem__(self, s, value):
        self.arr[s.start - self.row0: s.stop - self.row0] = value


def _check_strand(strand):
    if strand not in ("both", "minus", "plus"):
        print("Error, strand parameter of selectd_strand() should be choose from {'both', 'minus', 'plus'}",
              file=sys.stderr)
        sys.exit(1)


def _check_metric(metric):
    if metric not in METRICS:
        print("Error, unknown method {}".format(metric), file=sys.stderr)
        sys.exit(1)


# ---------------------------------------------------------------------------
# frequencies
# ---------------------------------------------------------------------------
def _read_genome(genome):
    return np.fromfile(genome, dtype=np.uint8)


def _profile_file(genome, pattern, strand, want):
    rank, world = ranks()
    if world > 1:
        return _profile_file_sharded(genome, pattern, strand, want, rank, world)
    text = _read_genome(genome)
    begin, end = engine.fasta_index(text)
    if begin.shape[0] == 0:
        _, _, dim = engine._lib.pattern_info(str(pattern))
        return None, 0, dim
    res = engine.profile_text(text, str(pattern), strand, want, begin, end)
    return res, int(begin.shape[0]), None


def _profile_file_sharded(genome, pattern, strand, want, rank, world):
    """Every rank profiles a contiguous, byte-balanced range of the records (the reference fans the
    sequences out to its workers, :868, :913, :967) and one NCCL all-gather gives every rank the
    whole (N, 4^k) matrix."""
    import torch.distributed as dist
    from . import sharding
    (key,) = want
    text = np.memmap(genome, dtype=np.uint8, mode="r") if os.path.getsize(genome) else np.zeros(0, np.uint8)
    begin, end = engine.fasta_index(text, threads=max(1, (os.cpu_count() or 1) // world))
    n = int(begin.shape[0])
    _, _, dim = engine._lib.pattern_info(str(pattern))
    if n == 0:
        return None, 0, dim
    cuts = sharding.record_cuts(end - begin, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    n_max = max(cuts[r + 1] - cuts[r] for r in range(world))
    dtype = torch.float64 if key == "freq64" else torch.float32
    device = engine.require_cuda()
    shard = torch.zeros((n_max, dim), dtype=dtype, device=device)
    if hi > lo:
        byte_lo, byte_hi = int(begin[lo]), int(end[hi - 1])
        res = engine.profile_text(np.ascontiguousarray(text[byte_lo:byte_hi]), str(pattern), strand, want,
                                  begin[lo:hi] - byte_lo, end[lo:hi] - byte_lo)
        shard[:hi - lo].copy_(res[key])
    gathered = torch.empty((world * n_max, dim), dtype=dtype, device=device)
    dist.all_gather_into_tensor(gathered, shard)
    full = torch.cat([gathered[r * n_max:r * n_max + cuts[r + 1] - cuts[r]] for r in range(world)], dim=0)
    return {key: full}, n, None


def compute_frequencies_device(genome, pattern, strand):
    """FASTA file -> (N, 4^k) float64 matrix (the --large None / scoop convention)."""
    res, n, dim = _profile_file(genome, pattern, strand, ("freq64",))
    if res is None:
        return np.zeros((0, dim), dtype=np.float64)
    return res["freq64"].cpu().numpy()


def _shared_tempdir(workdir):
    """A temp dir under workdir created by rank 0, same name on every rank."""
    rank, _ = ranks()
    return _broadcast_str(tempfile.mkdtemp(dir=workdir) if rank == 0 else None)


def compute_frequencies_memmap(genome, pattern, strand, workdir):
    """float32 memmap 'frequencies' in a temp dir under workdir (reference :879-916)."""
    rank, _ = ranks()
    folder = _shared_tempdir(workdir)
    freq_name = os.path.join(folder, "frequencies")
    res, n, dim = _profile_file(genome, pattern, strand, ("freq32",))
    if res is None:
        raise PhyloligoError("no FASTA record in {}".format(genome))
    f32 = res["freq32"]
    if rank == 0:
        frequencies = np.memmap(freq_name, dtype=np.float32, shape=tuple(f32.shape), mode="w+")
        frequencies[:] = f32.cpu().numpy()
        frequencies.flush()
        del frequencies
    _barrier()
    frequencies = np.memmap(freq_name, dtype=np.float32, shape=tuple(f32.shape), mode="r+")
    return frequencies, freq_name


def compute_frequencies_h5py(genome, pattern, strand, workdir):
    """HDF5 'frequencies_results' (dataset 'frequencies', float32) in a temp dir
    under workdir; returns (None, path) like the reference (:933-977)."""
    rank, _ = ranks()
    folder = _shared_tempdir(workdir)
    res, n, dim = _profile_file(genome, pattern, strand, ("freq32",))
    if res is None:
        raise PhyloligoError("no FASTA record in {}".format(genome))
    freq_name = os.path.join(folder, "frequencies_results")
    if rank == 0:
        io_formats.write_hdf5(freq_name, "frequencies", res["freq32"].cpu().numpy())
    _barrier()
    return None, freq_name


def compute_frequencies(mthdrun, large, genome, pattern, strand, distchunksize, threads_max, workdir):
    """Choose the output convention for the frequency stage (reference :980-997).
    distchunksize and threads_max are accepted and unused: the whole file is one
    GPU batch."""
    _check_strand(strand)
    freq_name = None
    if mthdrun == "scoop":
        frequencies = compute_frequencies_device(genome, pattern, strand)
    elif mthdrun == "joblib":
        if large == "memmap":
            frequencies, freq_name = compute_frequencies_memmap(genome, pattern, strand, workdir)
        elif large == "h5py":
            frequencies, freq_name = compute_frequencies_h5py(genome, pattern, strand, workdir)
        else:
            frequencies = compute_frequencies_device(genome, pattern, strand)
    else:
        print("Method {} is unknown".format(mthdrun), file=sys.stderr)
        sys.exit(1)
    return frequencies, freq_name


# ---------------------------------------------------------------------------
# distances
# ---------------------------------------------------------------------------
def compute_distances_device(frequencies, metric="Eucl"):
    """In-RAM float64 (N, N) matrix (the joblib/None and scoop conventions,
    reference :313-392).  Under torchrun every rank computes its block rows and rank 0
    receives them; the other ranks return None."""
    _check_metric(metric)
    device = engine.require_cuda()
    X = torch.from_numpy(np.ascontiguousarray(np.asarray(frequencies))).to(device)
    rank, world = ranks()
    if world == 1 or X.shape[0] < 2 * TILE * world:
        res = engine.distance_matrix_device(X, metric, torch.float64, symmetric=True).cpu().numpy()
        return res if rank == 0 else None
    import torch.distributed as dist
    from . import multigpu, sharding
    n = int(X.shape[0])
    P, aux, dim = engine.prepare(X, metric)
    job = multigpu.BlockRows(n, torch.float64, rank, world)
    job.compute(metric, P, aux, dim)
    full = torch.empty((n, n), dtype=torch.float64, device=device) if rank == 0 else None
    ops = []
    for i, (a, b) in enumerate(job.ranges):
        if b <= a:
            continue
        owner = sharding.range_owner(i, world)
        if rank == 0 and owner == 0:
            full[a:b].copy_(job.out_rows[i])
        elif rank == 0:
            ops.append(dist.P2POp(dist.irecv, full[a:b], owner))
        elif owner == rank:
            ops.append(dist.P2POp(dist.isend, job.out_rows[i], 0))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    res = full.cpu().numpy() if rank == 0 else None
    job.close()
    return res


def _stream_rows(sink_array, X, metric):
    """Fill the caller's (N, N) float32 array-like (memmap / HDF5 data region) with the matrix.
    One GPU: row panels, upper triangle + mirror when the matrix fits in HBM.  Under torchrun:
    every rank fills the rows it owns -- its paired block rows (upper triangle only, mirrored
    tiles stored into the owner's rows over NVLink) when they fit in HBM, else plain row panels."""
    rank, world = ranks()
    metric = _large_metric(metric)
    n = int(X.shape[0])

    # A finished panel is copied from the pinned buffer into the file mapping by a few threads
    # (numpy releases the GIL in the copy; one thread moves ~10 GB/s, the panels arrive at ~55 GB/s).
    from concurrent.futures import ThreadPoolExecutor
    workers = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    pool = ThreadPoolExecutor(workers) if workers > 1 else None

    def _put(a, b, host, r0):
        sink_array[a:b] = host[a - r0:b - r0]

    def sink(r0, r1, host):
        if pool is None or r1 - r0 < 4 * workers:
            sink_array[r0:r1] = host
            return
        step = -(-(r1 - r0) // workers)
        jobs = [pool.submit(_put, a, min(r1, a + step), host, r0) for a in range(r0, r1, step)]
        for j in jobs:
            j.result()

    try:
        _stream_rows_impl(sink, X, metric, n, rank, world)
    finally:
        if pool is not None:
            pool.shutdown()


def _stream_rows_impl(sink, X, metric, n, rank, world):
    if world == 1:
        engine.PanelStreamer(X, metric, torch.float32, PANEL_ROWS).run(sink)
        return
    from . import multigpu, sharding
    ranges = sharding.paired_row_ranges(n, world)
    rows_owned = sum(ranges[i][1] - ranges[i][0] for i in sharding.owned_ranges(ranges, rank, world))
    free, _ = torch.cuda.mem_get_info()
    fits = torch.tensor([1 if rows_owned * n * 4 < 0.6 * free and n >= 2 * TILE * world else 0], device=X.device)
    import torch.distributed as dist
    dist.all_reduce(fits, op=dist.ReduceOp.MIN)  # every rank must take the same path
    if int(fits.item()):
        P, aux, dim = engine.prepare(X, metric)
        job = multigpu.BlockRows(n, torch.float32, rank, world)
        job.compute(metric, P, aux, dim)
        panel = PANEL_ROWS or engine.auto_panel_rows(n, 4)
        pinned = [torch.empty((panel, n), dtype=torch.float32).pin_memory() for _ in range(2)]
        events, pending = [None, None], []
        k = 0
        for i in job.my_ranges:
            a, b = job.ranges[i]
            for r0 in range(a, b, panel):
                r1 = min(b, r0 + panel)
                slot = k & 1
                while pending and pending[0][0] == slot:
                    _, p0, p1 = pending.pop(0)
                    events[slot].synchronize()
                    sink(p0, p1, pinned[slot][:p1 - p0].numpy())
                pinned[slot][:r1 - r0].copy_(job.out_rows[i][r0 - a:r1 - a], non_blocking=True)
                events[slot] = torch.cuda.Event()
                events[slot].record()
                pending.append((slot, r0, r1))
                k += 1
        for slot, p0, p1 in pending:
            events[slot].synchronize()
            sink(p0, p1, pinned[slot][:p1 - p0].numpy())
        job.close()
    else:
        for i in sharding.owned_ranges(ranges, rank, world):
            a, b = ranges[i]
            if b > a:
                engine.PanelStreamer(X, metric, torch.float32, PANEL_ROWS, symmetric=False, rows=(a, b)).run(sink)


def compute_distances_memmap(frequencies, freq_name, output, metric="Eucl"):
    """Raw row-major float32 N x N file at `output` (reference :394-427)."""
    _check_metric(metric)
    device = engine.require_cuda()
    rank, _ = ranks()
    n = frequencies.shape[0]
    if rank == 0:
        np.memmap(output, dtype=np.float32, shape=(n, n), mode="w+").flush()
    _barrier()
    distances = np.memmap(output, dtype=np.float32, shape=(n, n), mode="r+")
    X = torch.from_numpy(np.ascontiguousarray(np.asarray(frequencies))).to(device)
    _stream_rows(distances, X, metric)
    distances.flush()
    del distances
    _barrier()
    if rank == 0:
        remove_folder(os.path.dirname(freq_name))


def compute_distances_h5py(freq_name, dist_name, metric="Eucl"):
    """HDF5 file at `dist_name`, dataset 'distances' (N, N) float32 (reference :480-534)."""
    _check_metric(metric)
    device = engine.require_cuda()
    rank, _ = ranks()
    freqs = io_formats.read_hdf5(freq_name, "frequencies")
    n = freqs.shape[0]
    X = torch.from_numpy(np.ascontiguousarray(freqs)).to(device)
    if rank == 0:
        io_formats.Hdf5DatasetWriter(dist_name, "distances", (n, n), np.float32).close()
    _barrier()
    with io_formats.Hdf5DatasetWriter.attach(dist_name, "distances") as writer:
        _stream_rows(writer.mm, X, metric)
    _barrier()
    if rank == 0:
        remove_folder(os.path.dirname(freq_name))


def compute_distances(mthdrun, large, frequencies, freq_name, out_file, dist, threads_max, freqchunksize, workdir):
    """Choose the output convention for the distance stage (reference :536-553)."""
    res = None
    if mthdrun == "joblib":
        if large == "memmap":
            compute_distances_memmap(frequencies, freq_name, out_file, metric=dist)
        elif large == "h5py":
            compute_distances_h5py(freq_name, out_file, metric=dist)
        else:
            res = compute_distances_device(frequencies, metric=dist)
    elif mthdrun == "scoop":
        res = compute_distances_device(frequencies, metric=dist)
    else:
        print("Error, method {} is not implemented for pairwise distances computation".format(mthdrun),
              file=sys.stderr)
    return res


# ---------------------------------------------------------------------------
# command line (flag surface of reference :1000-1034)
# ---------------------------------------------------------------------------
def get_cmd(argv=None):
    parser = argparse.ArgumentParser(description="Oligonucleotide-profile distances between contigs on B200 GPUs")
    parser.add_argument("-i", "--assembly", action="store", required=True, dest="genome",
                        help="multi-FASTA file of the assembly")
    parser.add_argument("-k", "--lgMot", action="store", dest="pattern", default=4, type=int,
                        help="k-mer length [default:%(default)d]")
    parser.add_argument("-s", "--strand", action="store", dest="strand", default="both",
                        choices=["both", "plus", "minus"], help="strand(s) profiled [default:%(default)s]")
    parser.add_argument("-d", "--distance", action="store", dest="dist", default="Eucl",
                        choices=["Eucl", "JSD", "KT", "BC", "SC"],
                        help="Eucl: Euclidean, JSD: Jensen-Shannon divergence, KT: Kendall's tau, "
                             "BC: Bray-Curtis, SC: Spearman correlation [default:%(default)s]")
    parser.add_argument("--freq-chunk-size", action="store", dest="freqchunksize", type=int, default=250,
                        help="accepted for compatibility (scoop chunking of the reference)")
    parser.add_argument("--dist-chunk-size", action="store", dest="distchunksize", type=int, default=250,
                        help="accepted for compatibility (scoop chunking of the reference)")
    parser.add_argument("--method", action="store", choices=["scoop", "joblib"], default="joblib", dest="mthdrun",
                        required=True, help="output convention of the reference back-end; the work runs on the GPU")
    parser.add_argument("--large", action="store", dest="large", choices=["None", "memmap", "h5py"], default="None",
                        help="stream large results to a float32 memmap or HDF5 file")
    parser.add_argument("-c", "--cpu", action="store", dest="threads_max", type=int, default=4,
                        help="accepted for compatibility [default:%(default)d]")
    parser.add_argument("-o", "--out", action="store", dest="out_file", default="phyloligo.out",
                        help="output file [default:%(default)s]")
    parser.add_argument("-q", "--outfreq", action="store", dest="out_freq_file",
                        help="also write the frequency matrix to this file")
    parser.add_argument("-w", "--workdir", action="store", dest="workdir", default=".", help="working directory")
    parser.add_argument("-p", "--pattern", action="store", dest="pattern", default="1111",
                        help="spaced-word pattern of 1s and 0s, e.g. '100101001' [default:'1111']")
    params = parser.parse_args(argv)
    params.workdir = os.path.abspath(params.workdir)
    return params


def main(argv=None):
    params = get_cmd(argv)
    if type(params.pattern) == int:  # -k given last: k-mer -> pattern without joker
        params.pattern = "1" * params.pattern

    rank, world = ranks()
    say = print if rank == 0 else (lambda *a, **k: None)
    say("Using pattern {}".format(params.pattern))
    if rank == 0 and not os.path.isdir(params.workdir):
        os.makedirs(params.workdir)
    _barrier()

    import time
    verbose = os.environ.get("PO_VERBOSE") == "1" and rank == 0  # stage timings on stderr
    t0 = time.perf_counter()
    say("Computing frequencies")
    frequencies, freq_name = compute_frequencies(params.mthdrun, params.large, params.genome, params.pattern,
                                                 params.strand, params.distchunksize, params.threads_max,
                                                 params.workdir)
    freq_for_q = None
    if params.out_freq_file:
        # the distance stage removes the temp dir, so fetch what -q needs first
        if frequencies is None:
            freq_for_q = io_formats.read_hdf5(freq_name, "frequencies")
        else:
            freq_for_q = np.array(frequencies)

    t1 = time.perf_counter()
    say("Computing Pairwise distances")
    res = compute_distances(params.mthdrun, params.large, frequencies, freq_name, params.out_file, params.dist,
                            params.threads_max, params.freqchunksize, params.workdir)

    t2 = time.perf_counter()
    if params.out_freq_file and rank == 0:
        say("Writing frequency matrix")
        io_formats.savetxt(params.out_freq_file, freq_for_q)

    if not (params.mthdrun == "joblib" and params.large != "None") and rank == 0:
        say("Writing distance matrix")
        io_formats.savetxt(params.out_file, res)
    if verbose:
        print("phyloligo_b200: frequencies %.3f s, distances %.3f s, text output %.3f s"
              % (t1 - t0, t2 - t1, time.perf_counter() - t2), file=sys.stderr)
    if world > 1:
        import torch.distributed as dist
        _barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    main()
    sys.exit(0)

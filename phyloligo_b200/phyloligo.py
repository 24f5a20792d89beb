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

from . import engine, hostsink, io_formats, phylodist
from ._lib import PhyloligoError, METRICS, TILE, FLAG_MIRROR, FLAG_SKIP_LOWER

PANEL_ROWS = None  # rows per computed panel of the streamed outputs; None = 4096
RESIDENT_FRACTION = 0.6  # the matrix (or a rank's rows of it) stays on the device when it fits in this share of the free HBM


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
_GENOME_INDEX = {}  # abspath -> (size, mtime_ns, begin, end): main() indexes the file while the CUDA context comes up


def _map_genome(genome):
    """The FASTA file mapped read-only (uint8 array; nothing is read until it is touched)."""
    if os.path.getsize(genome) == 0:
        return np.zeros(0, np.uint8)
    return np.memmap(genome, dtype=np.uint8, mode="r")


def _genome_index(genome, threads=0):
    """(begin, end) of every record of the file (engine.fasta_index over the mapping), cached."""
    key = os.path.abspath(genome)
    st = os.stat(genome)
    hit = _GENOME_INDEX.get(key)
    if hit is not None and hit[0] == st.st_size and hit[1] == st.st_mtime_ns:
        return hit[2], hit[3]
    begin, end = engine.fasta_index(_map_genome(genome), threads=threads)
    _GENOME_INDEX.clear()
    _GENOME_INDEX[key] = (st.st_size, st.st_mtime_ns, begin, end)
    return begin, end


def _profile_file(genome, pattern, strand, want):
    rank, world = ranks()
    if world > 1:
        return _profile_file_sharded(genome, pattern, strand, want, rank, world)
    begin, end = _genome_index(genome)
    if begin.shape[0] == 0:
        _, _, dim = engine._lib.pattern_info(str(pattern))
        return None, 0, dim
    device = engine.require_cuda()
    lo, hi = int(begin[0]), int(end[-1])
    d_text = engine.file_to_device(genome, lo, hi, device)  # header of the first record and before: not needed
    d_begin = torch.from_numpy(begin - lo).to(device)
    d_end = torch.from_numpy(end - lo).to(device)
    res = engine.profile_device(d_text, d_begin, d_end, str(pattern), strand, want)
    return res, int(begin.shape[0]), None


def _profile_file_sharded(genome, pattern, strand, want, rank, world):
    """Every rank profiles a contiguous, byte-balanced range of the records (the reference fans the
    sequences out to its workers, :868, :913, :967) and one NCCL all-gather gives every rank the
    whole (N, 4^k) matrix."""
    import torch.distributed as dist
    from . import sharding
    (key,) = want
    begin, end = _genome_index(genome, threads=max(1, (os.cpu_count() or 1) // world))
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
        d_text = engine.file_to_device(genome, byte_lo, byte_hi, device)
        d_begin = torch.from_numpy(begin[lo:hi] - byte_lo).to(device)
        d_end = torch.from_numpy(end[lo:hi] - byte_lo).to(device)
        res = engine.profile_device(d_text, d_begin, d_end, str(pattern), strand, want)
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
    if world == 1:
        n = int(X.shape[0])
        res = np.empty((n, n), dtype=np.float64)
        if metric == "KT" and X.shape[1] < 2:
            res.fill(1.0)  # kendall() with no element pair: distance 0 -> KT = 1
        else:
            _fill_host_matrix(res, X, metric)  # panels when the float64 matrix does not fit in HBM
        return res
    if X.shape[0] < 2 * TILE * world:
        res = engine.distance_matrix_device(X, metric, torch.float64, symmetric=True).cpu().numpy()
        return res if rank == 0 else None
    # Every rank ships the rows it owns into one scratch file in shared memory (the same sink as the --large
    # outputs: nothing the size of the matrix is ever allocated on a device or sent through rank 0's GPU) and
    # rank 0 reads the matrix back as the in-RAM float64 array of this convention.
    n = int(X.shape[0])
    scratch_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = None
    if rank == 0:
        fd, path = tempfile.mkstemp(prefix="phyloligo_b200_", suffix=".f64", dir=scratch_dir)
        os.close(fd)
    path = _broadcast_str(path)
    fm = _open_file_matrix(path, n, dtype=np.float64)
    try:
        _fill_host_matrix(fm.array, X, metric)
    finally:
        fm.close()
    _barrier()
    res = None
    if rank == 0:
        res = np.fromfile(path, dtype=np.float64).reshape(n, n)
        os.unlink(path)
    return res


_PREPARED = {}  # abspath -> hostsink.FileMatrix created early by main() (its pages are already being instantiated)


def _my_row_ranges(n, rank, world):
    """Row ranges this rank fills, in the order it fills them."""
    if world == 1:
        return [(0, n)]
    from . import sharding
    ranges = sharding.paired_row_ranges(n, world)
    return [ranges[i] for i in sharding.owned_ranges(ranges, rank, world) if ranges[i][1] > ranges[i][0]]


def _open_file_matrix(path, n, offset=0, create=True, fresh=None, dtype=np.float32):
    """The output region as a hostsink.FileMatrix whose pages (of this rank's rows) are being instantiated.
    Rank 0 creates the file, the others attach after the barrier."""
    rank, world = ranks()
    fm = _PREPARED.pop(os.path.abspath(path), None)
    if fm is not None and (fm.rows, fm.cols, fm.offset) == (n, n, offset):
        return fm
    if fm is not None:
        fm.close()
    if rank == 0:
        fm = hostsink.FileMatrix(path, n, n, dtype, offset, create=create)
    _barrier()
    if rank != 0:
        fm = hostsink.FileMatrix(path, n, n, dtype, offset, create=False)
    # rank 0 knows whether the file's pages exist already
    fresh = fresh if fresh is not None else _broadcast_str(fm.fresh if rank == 0 else None)
    threads = hostsink.host_threads(world)
    fm.warm(_my_row_ranges(n, rank, world), threads=max(1, min(4, threads // 2)), fresh=bool(fresh))
    return fm


def _fill_host_matrix(dest, X, metric):
    """Fill the caller's (N, N) host array `dest` -- the mapping of the raw memmap file or of the HDF5
    data region (float32), or the in-RAM result of the --large None conventions (float64) -- with the
    matrix.  One GPU: row panels, upper triangle + mirror when the matrix fits in HBM; every finished
    panel leaves through hostsink.RowShipper (DMA into a pinned ring, host threads into `dest`)
    while the next one computes.  Under torchrun every rank fills the rows it owns -- its paired
    block rows (upper triangle only, mirrored tiles stored into the owner's rows over NVLink) when
    they fit in HBM, else plain row panels."""
    rank, world = ranks()
    n = int(X.shape[0])
    if n == 0:
        return
    out_dtype = torch.float32 if dest.dtype == np.float32 else torch.float64
    esize = dest.itemsize
    threads = hostsink.host_threads(world)
    # ring slots of at most 64 MB (page-locking costs ~0.5 s per GB: they join the ring one by one), no larger than the job needs
    slot = int(min(64 << 20, max(1 << 20, n * esize, n * n * esize // 2)))
    shipper = hostsink.RowShipper(dest, slot_bytes=slot, copy_threads=max(1, min(8, threads - min(4, threads // 2))))
    panel = PANEL_ROWS or 4096
    panel = max(TILE, (int(panel) // TILE) * TILE)
    try:
        if world == 1:
            free, _ = torch.cuda.mem_get_info()
            P, aux, dim = engine.prepare(X, metric)
            if n * n * esize < RESIDENT_FRACTION * free:
                full = torch.empty((n, n), dtype=out_dtype, device=X.device)
                for r0 in range(0, n, panel):
                    r1 = min(n, r0 + panel)
                    engine.distance_block(metric, P, aux, dim, r0, r1, 0, n, full, 0, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
                    shipper.ship(full[r0:r1], r0, 0)  # rows [r0, r1) are final: earlier panels mirrored into them
            else:
                _fill_panels(shipper, metric, P, aux, dim, n, [(0, n)], panel, X.device, out_dtype)
        else:
            from . import multigpu, sharding
            import torch.distributed as dist
            ranges = sharding.paired_row_ranges(n, world)
            rows_owned = sum(b - a for a, b in _my_row_ranges(n, rank, world))
            free, _ = torch.cuda.mem_get_info()
            fits = torch.tensor([1 if rows_owned * n * esize < RESIDENT_FRACTION * free and n >= 2 * TILE * world else 0], device=X.device)
            dist.all_reduce(fits, op=dist.ReduceOp.MIN)  # every rank must take the same path
            P, aux, dim = engine.prepare(X, metric)
            if int(fits.item()):
                job = multigpu.BlockRows(n, out_dtype, rank, world)
                job.compute(metric, P, aux, dim, ship=shipper.ship)
                shipper.finish()
                shipper = None
                job.close()
            else:
                _fill_panels(shipper, metric, P, aux, dim, n, _my_row_ranges(n, rank, world), panel, X.device, out_dtype)
    finally:
        if shipper is not None:
            shipper.finish()


def _fill_panels(shipper, metric, P, aux, dim, n, row_ranges, panel, device, out_dtype=torch.float32):
    """Plain block rows (every entry computed) through two device panel buffers."""
    bufs = [torch.empty((panel, n), dtype=out_dtype, device=device) for _ in range(2)]
    busy = [None, None]
    k = 0
    for a, b in row_ranges:
        for r0 in range(a, b, panel):
            r1 = min(b, r0 + panel)
            slot = k & 1
            if busy[slot] is not None:
                torch.cuda.current_stream().wait_event(busy[slot])  # its rows have left the device
            engine.distance_block(metric, P, aux, dim, r0, r1, 0, n, bufs[slot], r0, 0, 0)
            busy[slot] = shipper.ship(bufs[slot][:r1 - r0], r0, 0)
            k += 1


def _profiles_to_device(frequencies):
    device = engine.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(frequencies))).to(device)


def compute_distances_memmap(frequencies, freq_name, output, metric="Eucl"):
    """Raw row-major float32 N x N file at `output` (reference :394-427)."""
    _check_metric(metric)
    rank, _ = ranks()
    n = int(frequencies.shape[0])
    fm = _open_file_matrix(output, n)
    try:
        _fill_host_matrix(fm.array, _profiles_to_device(frequencies), _large_metric(metric))
    finally:
        fm.close()
    _barrier()
    if rank == 0:
        remove_folder(os.path.dirname(freq_name))


def compute_distances_h5py(freq_name, dist_name, metric="Eucl"):
    """HDF5 file at `dist_name`, dataset 'distances' (N, N) float32 (reference :480-534)."""
    _check_metric(metric)
    rank, _ = ranks()
    freqs = io_formats.read_hdf5(freq_name, "frequencies")
    n = int(freqs.shape[0])
    fm = _PREPARED.get(os.path.abspath(dist_name))
    if fm is None:
        if rank == 0:
            io_formats.Hdf5DatasetWriter(dist_name, "distances", (n, n), np.float32).close()
        _barrier()
        _, _, offset = io_formats.dataset_location(dist_name, "distances")
    else:
        offset = fm.offset
    fm = _open_file_matrix(dist_name, n, offset, create=False, fresh=True)  # the writer has just truncated the file to size
    try:
        _fill_host_matrix(fm.array, _profiles_to_device(freqs), _large_metric(metric))
    finally:
        fm.close()
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


def _warm_cuda():
    try:
        torch.cuda.init()
        torch.empty(1, device="cuda")
        engine._lib.load()
    except Exception:  # the stages report the missing device / library themselves
        pass


def _prepare_outputs(params):
    """Index the assembly and, when the distance stage will fill a file, create it now."""
    if not os.path.isfile(params.genome) or params.strand not in ("both", "minus", "plus"):
        return
    begin, _ = _genome_index(params.genome)
    n = int(begin.shape[0])
    if n == 0 or params.mthdrun != "joblib" or params.dist not in METRICS:
        return
    out = os.path.abspath(params.out_file)
    threads = hostsink.host_threads(1)
    if params.large == "memmap":
        fm = hostsink.FileMatrix(params.out_file, n, n, np.float32, 0, create=True)
    elif params.large == "h5py":
        writer = io_formats.Hdf5DatasetWriter(params.out_file, "distances", (n, n), np.float32)
        offset = writer.data_off
        writer.close()
        fm = hostsink.FileMatrix(params.out_file, n, n, np.float32, offset, create=False)
        fm.fresh = True
    else:
        return
    fm.warm([(0, n)], threads=max(1, min(4, threads // 2)))
    _PREPARED[out] = fm


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

    import threading
    import time
    verbose = os.environ.get("PO_VERBOSE") == "1" and rank == 0  # stage timings on stderr
    t0 = time.perf_counter()
    if world == 1:
        # While the CUDA context comes up (seconds on a fresh process) the host indexes the FASTA file
        # and, for the --large outputs, creates the N x N file and starts instantiating its pages --
        # the slowest part of the whole run on a fresh file (hostsink.py).
        warm = threading.Thread(target=_warm_cuda, daemon=True)
        warm.start()
        try:
            _prepare_outputs(params)
        finally:
            warm.join()
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

#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: run the reference's UNMODIFIED command-line script in this container.

    python oracle/run_reference_cli.py -i asm.fasta -k 4 -d JSD --method joblib -o out.mat [...]

/root/reference/phylopackage/bin/phyloligo.py cannot be imported as it is: `scoop`, `Bio`
(Biopython) and `h5py` are not installed and `sklearn.externals.joblib` no longer exists.  This
runner injects minimal stand-ins for exactly those third-party modules into ``sys.modules`` and then
executes the script with ``runpy`` -- every line of the reference itself runs as written
("Full-CLI oracle").  The stand-ins:

  scoop.futures              map = builtin map
  sklearn.externals.joblib   the installed joblib (Parallel, delayed, dump, load)
  Bio.SeqIO.parse            a FASTA reader with Biopython's record rules: a record starts at a line
                             beginning with '>', text before the first '>' is ignored, the sequence is
                             every later line up to the next header with whitespace removed
  Bio.Seq.Seq                str subclass with reverse_complement() (IUPAC DNA table, both cases)
  Bio.Cluster.distancematrix the restated kendall() of the C Clustering Library (dist="k") -- the only
                             stand-in that carries arithmetic of its own (oracle/phylo_oracle.py, KT:
                             parity unpinned, see there)
  h5py                       absent: the --large h5py mode cannot be run here

  matplotlib, hdbscan        empty modules (imported at the top of bin/phyloselect.py; plotting, t-SNE and
                             HDBSCAN are not exercised: `-m kmedoids` without `-t`)
  Bio.SeqIO.write            FASTA writer as Biopython's: ">" + title line, sequence wrapped at 60 columns

``--script phyloselect.py`` (first argument) runs another script of phylopackage/bin the same way.
Nothing of the reference is copied; it is read where it lies.  Used by tests/golden/make_cli_golden.py
and make_select_cli_golden.py to produce the committed end-to-end fixtures (the GPU box has no
/root/reference).
"""
import os
import runpy
import sys
import types

REFERENCE_ROOT = os.environ.get("PHYLOLIGO_REFERENCE", "/root/reference")
SCRIPT = os.path.join(REFERENCE_ROOT, "phylopackage", "bin", "phyloligo.py")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_COMP = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn")


class Seq(str):
    def reverse_complement(self):
        return Seq(str(self).translate(_COMP)[::-1])

    def upper(self):
        return Seq(str.upper(self))


class _Record:
    def __init__(self, title, seq):
        self.id = title.split()[0] if title.split() else ""
        self.description = title
        self.seq = Seq(seq)


def _parse(handle, fmt):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, (str, bytes, os.PathLike)):
        handle = open(handle)
        close = True
    try:
        title, parts = None, []
        for line in handle:
            if line.startswith(">"):
                if title is not None:
                    yield _Record(title, "".join(parts))
                title, parts = line[1:].rstrip(), []
            elif title is not None:
                parts.append("".join(line.split()))
        if title is not None:
            yield _Record(title, "".join(parts))
    finally:
        if close:
            handle.close()


def install_shims():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    scoop = types.ModuleType("scoop")
    scoop.futures = types.ModuleType("scoop.futures")
    scoop.futures.map = map
    sys.modules.setdefault("scoop", scoop)
    sys.modules.setdefault("scoop.futures", scoop.futures)

    import joblib
    import sklearn
    jl = joblib
    if os.environ.get("PO_REF_JOBLIB_THREADS") == "1":
        # the --large memmap workers take an np.memmap argument that loky's worker processes cannot
        # un-serialize for functions of a runpy __main__: same calls, thread workers
        jl = types.ModuleType("joblib_threads")
        jl.delayed, jl.dump, jl.load = joblib.delayed, joblib.dump, joblib.load

        def Parallel(*args, **kwargs):
            kwargs.setdefault("backend", "threading")
            return joblib.Parallel(*args, **kwargs)

        jl.Parallel = Parallel
    ext = types.ModuleType("sklearn.externals")
    ext.joblib = jl
    sklearn.externals = ext
    sys.modules["sklearn.externals"] = ext
    sys.modules["sklearn.externals.joblib"] = jl

    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse = _parse

    def _write(records, handle, fmt):
        assert fmt == "fasta"
        count = 0
        for rec in records:
            handle.write(">%s\n" % rec.description)
            seq = str(rec.seq)
            for p in range(0, len(seq), 60):
                handle.write(seq[p:p + 60] + "\n")
            count += 1
        return count

    seqio.write = _write
    seqmod = types.ModuleType("Bio.Seq")
    seqmod.Seq = Seq
    cluster = types.ModuleType("Bio.Cluster")

    def distancematrix(data, dist="e", **kwargs):
        from oracle import phylo_oracle as po
        assert dist == "k", "only the Kendall distance is used by the reference (core/phylodist.py:74)"
        rows = [list(r) for r in data]
        return [[]] + [[1.0 - po.KT(rows[i], rows[j]) for j in range(i)] for i in range(1, len(rows))]

    cluster.distancematrix = distancematrix
    bio.SeqIO, bio.Seq, bio.Cluster = seqio, seqmod, cluster
    for name, mod in (("Bio", bio), ("Bio.SeqIO", seqio), ("Bio.Seq", seqmod), ("Bio.Cluster", cluster)):
        sys.modules.setdefault(name, mod)
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")  # imported at the top of the script, used by --large h5py only
    for name in ("matplotlib", "matplotlib.pyplot", "hdbscan"):  # top-level imports of bin/phyloselect.py
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                mod = types.ModuleType(name)
                mod.use = lambda *a, **k: None
                sys.modules[name] = mod
    if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)  # `from phylopackage import phylodist`


def run(argv):
    """Execute a reference script (default bin/phyloligo.py; ``--script NAME`` first selects another one of
    phylopackage/bin) with the given command-line arguments (list of str)."""
    argv = list(argv)
    script = SCRIPT
    if argv[:1] == ["--script"]:
        script = os.path.join(REFERENCE_ROOT, "phylopackage", "bin", argv[1])
        argv = argv[2:]
    if not os.path.isfile(script):
        raise RuntimeError("reference checkout not mounted at %s" % REFERENCE_ROOT)
    install_shims()
    old = sys.argv
    sys.argv = [script] + argv
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit as exc:
        if exc.code not in (0, None):
            raise
    finally:
        sys.argv = old


if __name__ == "__main__":
    run(sys.argv[1:])

"""Row N2 of SURVEY.md section 8f on the CPU: the snapshot / mesh writers of csrc/host/io.c (the bigfile on-disk format restated)
against the directories the compiled reference writes with its own libfastpmio/io.c + depends/bigfile (oracle/_ref): same files,
byte for byte; and each side reads what the other wrote."""
import ctypes as C
import filecmp
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COSMO = np.array([0.307494, 0.6774, 0.0, 3.046, 0, -1.0, 0.0])      # Omega_m, h, T_cmb, N_eff, N_nu, w0, wa (the oracle session's defaults)


class IoColumn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype_out", C.c_char_p), ("dtype", C.c_char_p), ("nmemb", C.c_int), ("data", C.c_void_p), ("on_device", C.c_int)]


class IoMeta(C.Structure):
    _fields_ = [("q_strides", C.c_int64 * 3), ("q_scale", C.c_double * 3), ("q_shift", C.c_double * 3), ("q_size", C.c_int64),
                ("a_x", C.c_double), ("a_v", C.c_double), ("M0", C.c_double)]


class IoHeader(C.Structure):
    _fields_ = [("NC", C.c_int64)] + [(n, C.c_double) for n in ("BoxSize", "ScalingFactor", "GrowthFactor", "GrowthRate", "HubbleE", "RSDFactor",
                                                                "Omega_cdm", "OmegaM", "OmegaLambda", "HubbleParam")] + \
               [("version", C.c_char_p), ("TotNumPart", C.c_uint64 * 6), ("MassTable", C.c_double * 6)]


@pytest.fixture(scope="module")
def lib():
    path = os.path.join(ROOT, "fastpm_b200", "libfastpm_b200.so")
    if not os.path.exists(path):
        pytest.skip("libfastpm_b200.so not built")
    return C.CDLL(path)


def _files(top):
    out = []
    for d, _, fs in os.walk(top):
        out += [os.path.relpath(os.path.join(d, f), top) for f in fs]
    return sorted(out)


def _attrs(path):
    out = {}
    for line in open(path):
        name, dtype, nmemb, hexdata = line.split()[:4]
        out[name] = (dtype, int(nmemb), bytes.fromhex(hexdata), line)
    return out


@pytest.fixture(scope="module")
def run(ref_mod, pk_text, tmp_path_factory):
    """A small reference run whose final state the reference writes as a snapshot."""
    tmp = tmp_path_factory.mktemp("snap")
    nc, L = 8, 32.0
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0, compute_potential=True)
    dk, _, _ = s.ic_deltak(7, pk_text)
    s.setup_lpt(dk, 0.1)
    s.evolve(np.linspace(0.1, 0.9, 3))
    ref_dir = str(tmp / "ref_0.9000")
    s.write_snapshot(ref_dir)
    x, v = s.snapshot_particles()
    p = s.get_particles()
    mesh_dir = str(tmp / "ref_dk")
    s.write_complex(dk, mesh_dir, "LinearDensityK", which=1)
    dk_c = s.complex_view(dk, which=1)
    yield dict(session=s, ref_dir=ref_dir, x=x, v=v, p=p, nc=nc, L=L, tmp=tmp, mesh_dir=mesh_dir, dk=dk_c, a=0.9)
    s.close()


def _potential(ref_dir, n):
    return np.fromfile(os.path.join(ref_dir, "1", "Potential", "000000"), dtype=np.float32, count=n)


def test_snapshot_directory_is_byte_identical(lib, run):
    nc, L, n = run["nc"], run["L"], len(run["x"])
    mine = str(run["tmp"] / "mine_0.9000")
    M0 = float(run["p"]["meta"][2])
    h = IoHeader()
    lib.fastpm_b200_host_io_header.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_void_p]
    lib.fastpm_b200_host_io_header(COSMO.ctypes.data, 0, nc, L, run["a"], M0, n, C.byref(h))
    lib.fastpm_b200_io_write_header.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    assert lib.fastpm_b200_io_write_header(mine.encode(), C.byref(h), 1) == 0
    # the columns of the store in the order of io.c:392-421; the potential column comes from the reference file (the unit
    # conversion of solver.c:647-702 is not what this test is about)
    x, v, p = run["x"], run["v"], run["p"]
    pot = _potential(run["ref_dir"], n)
    ids = p["id"]
    keep = [x, v, p["dx1"], p["dx2"], ids, pot]
    cols = (IoColumn * 6)(
        IoColumn(b"Position", b"f4", b"f8", 3, x.ctypes.data, 0), IoColumn(b"DX1", b"f4", b"f4", 3, p["dx1"].ctypes.data, 0),
        IoColumn(b"DX2", b"f4", b"f4", 3, p["dx2"].ctypes.data, 0), IoColumn(b"Velocity", b"f4", b"f4", 3, v.ctypes.data, 0),
        IoColumn(b"ID", b"i8", b"i8", 1, ids.ctypes.data, 0), IoColumn(b"Potential", b"f4", b"f4", 1, pot.ctypes.data, 0))
    m = IoMeta((C.c_int64 * 3)(nc * nc, nc, 1), (C.c_double * 3)(L / nc, L / nc, L / nc), (C.c_double * 3)(0, 0, 0), nc ** 3, run["a"], run["a"], M0)
    lib.fastpm_b200_io_write_columns.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int]
    assert lib.fastpm_b200_io_write_columns(mine.encode(), b"1", cols, 6, n, C.byref(m), 1) == 0
    del keep
    assert _files(mine) == _files(run["ref_dir"])
    for f in _files(mine):
        if f == os.path.join("Header", "attr-v2"):
            continue
        assert filecmp.cmp(os.path.join(mine, f), os.path.join(run["ref_dir"], f), shallow=False), f
    # the header attributes: same names, types and order; bytes equal except the growth numbers (our QAG / RKF45 against GSL:
    # 2e-7, tests/test_cpu_oracle_and_host.py) and the library version string
    a, b = _attrs(os.path.join(mine, "Header", "attr-v2")), _attrs(os.path.join(run["ref_dir"], "Header", "attr-v2"))
    assert list(a) == list(b)
    for name in a:
        assert a[name][:2] == b[name][:2] or name == "LibFastPMVersion", name
        if name in ("GrowthFactor", "GrowthRate", "HubbleE", "RSDFactor"):
            np.testing.assert_allclose(np.frombuffer(a[name][2], dtype=np.float64), np.frombuffer(b[name][2], dtype=np.float64), rtol=2e-7)
        elif name != "LibFastPMVersion":
            assert a[name][3] == b[name][3], name


def test_appended_catalog_is_byte_identical(lib, run, tmp_path):
    """fastpm_store_write in append mode (io.c:522-537): every column block grows by big_block_mpi_grow_simple and the new items land
    behind the old end; appending to a catalog that does not exist yet creates empty blocks first.  Same files, byte for byte, as the
    reference's, after write + append and after append + append."""
    s, n, nc, L = run["session"], len(run["x"]), run["nc"], run["L"]
    x, v, p = run["x"], run["v"], run["p"]
    pot = _potential(run["ref_dir"], n)
    ids = p["id"]
    keep = [x, v, p["dx1"], p["dx2"], ids, pot]
    cols = (IoColumn * 6)(
        IoColumn(b"Position", b"f4", b"f8", 3, x.ctypes.data, 0), IoColumn(b"DX1", b"f4", b"f4", 3, p["dx1"].ctypes.data, 0),
        IoColumn(b"DX2", b"f4", b"f4", 3, p["dx2"].ctypes.data, 0), IoColumn(b"Velocity", b"f4", b"f4", 3, v.ctypes.data, 0),
        IoColumn(b"ID", b"i8", b"i8", 1, ids.ctypes.data, 0), IoColumn(b"Potential", b"f4", b"f4", 1, pot.ctypes.data, 0))
    m = IoMeta((C.c_int64 * 3)(nc * nc, nc, 1), (C.c_double * 3)(L / nc, L / nc, L / nc), (C.c_double * 3)(0, 0, 0), nc ** 3, run["a"], run["a"],
               float(p["meta"][2]))
    lib.fastpm_b200_io_write_columns.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int]
    lib.fastpm_b200_io_append_columns.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64, C.c_int]
    for first in ("w", "a"):
        ref_dir, mine = str(tmp_path / ("ref_" + first)), str(tmp_path / ("mine_" + first))
        if first == "w":
            s.write_snapshot(ref_dir)
            assert lib.fastpm_b200_io_write_columns(mine.encode(), b"1", cols, 6, n, C.byref(m), 1) == 0
        else:
            s.append_snapshot(ref_dir)
            assert lib.fastpm_b200_io_append_columns(mine.encode(), b"1", cols, 6, n, 1) == 0
        s.append_snapshot(ref_dir)
        assert lib.fastpm_b200_io_append_columns(mine.encode(), b"1", cols, 6, n, 1) == 0
        theirs = [f for f in _files(ref_dir) if f.startswith("1" + os.sep)]
        ours = [f for f in _files(mine) if f.startswith("1" + os.sep)]
        assert ours == theirs, (first, ours, theirs)
        for f in theirs:
            assert filecmp.cmp(os.path.join(mine, f), os.path.join(ref_dir, f), shallow=False), (first, f)
        # two files of n items each per column now
        head = open(os.path.join(mine, "1", "ID", "header")).read()
        assert "NFILE: 2" in head and head.count(": %d :" % n) == 2, head
    del keep


def test_reference_reads_our_snapshot_and_we_read_theirs(lib, run, ref_mod, pk_text):
    nc, L, n = run["nc"], run["L"], len(run["x"])
    mine = str(run["tmp"] / "mine_0.9000")
    assert os.path.isdir(mine), "run after test_snapshot_directory_is_byte_identical"
    # the reference's fastpm_store_read + read_snapshot_header on our directory
    s = ref_mod.Session(nc=nc, boxsize=L, pm_nc_factor=2, force_mode="cola", growth_mode="LCDM", np_alloc_factor=2.0, compute_potential=True)
    assert s.read_snapshot(mine) == run["a"]
    q = s.get_particles()
    s.close()
    assert np.array_equal(q["x"], run["x"].astype(np.float32).astype(np.float64))
    assert np.array_equal(q["v"], run["v"]) and np.array_equal(q["id"], run["p"]["id"]) and np.array_equal(q["dx1"], run["p"]["dx1"])
    assert np.array_equal(q["meta"][:2], [run["a"], run["a"]])
    # our reader on the reference's directory
    x = np.zeros((n + 5, 3)); v = np.zeros((n + 5, 3), dtype=np.float32); ids = np.zeros(n + 5, dtype=np.uint64)
    cols = (IoColumn * 3)(IoColumn(b"Position", b"f4", b"f8", 3, x.ctypes.data, 0), IoColumn(b"Velocity", b"f4", b"f4", 3, v.ctypes.data, 0),
                          IoColumn(b"ID", b"i8", b"i8", 1, ids.ctypes.data, 0))
    m = IoMeta()
    cap = C.c_int64(n + 5)
    lib.fastpm_b200_io_read_columns.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    assert lib.fastpm_b200_io_read_columns(run["ref_dir"].encode(), b"1", cols, 3, C.byref(cap), C.byref(m), 1) == 0
    assert cap.value == n and m.q_size == nc ** 3 and m.a_x == run["a"] and list(m.q_strides) == [nc * nc, nc, 1]
    assert np.array_equal(x[:n], run["x"].astype(np.float32).astype(np.float64)) and np.array_equal(v[:n], run["v"]) and np.array_equal(ids[:n], run["p"]["id"])
    cap = C.c_int64(n - 1)                              # more items than the store can hold: refused like io.c:495-498
    assert lib.fastpm_b200_io_read_columns(run["ref_dir"].encode(), b"1", cols, 3, C.byref(cap), C.byref(m), 1) != 0


def test_mesh_block_is_byte_identical_and_round_trips(lib, run):
    n = run["nc"]                                      # the LPT mesh of this run is nc
    hc = n // 2 + 1
    pitch_c = ((hc + 15) // 16) * 16
    rows = np.zeros((n, n, pitch_c), dtype=np.complex64)           # device layout [ky][kx][kz], padded rows
    rows[:, :, :hc] = np.transpose(run["dk"], (1, 0, 2))
    mine = str(run["tmp"] / "mine_dk")
    lib.fastpm_b200_io_write_complex_rows.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    assert lib.fastpm_b200_io_write_complex_rows(mine.encode(), b"LinearDensityK", n, run["L"], 0, n, rows.ctypes.data, pitch_c, 1, 1) == 0
    assert _files(mine) == _files(run["mesh_dir"])
    for f in _files(mine):
        assert filecmp.cmp(os.path.join(mine, f), os.path.join(run["mesh_dir"], f), shallow=False), f
    back = np.zeros_like(rows)
    lib.fastpm_b200_io_read_complex_rows.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    assert lib.fastpm_b200_io_read_complex_rows(run["mesh_dir"].encode(), b"LinearDensityK", n, 0, n, back.ctypes.data, pitch_c) == 0
    assert np.array_equal(back, rows)
    # two writers (two slabs, two catalog slices): tests/mp_worker.py, world_size 2 over gloo


def test_argsort_by_id_is_a_stable_radix_sort(lib):
    rng = np.random.default_rng(2)
    key = rng.integers(0, 1 << 40, size=20000, dtype=np.uint64)
    key[::7] = key[3]                                   # duplicates: stability decides
    perm = np.zeros(len(key), dtype=np.uint64)
    lib.fastpm_b200_io_argsort_u64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.fastpm_b200_io_argsort_u64(key.ctypes.data, len(key), perm.ctypes.data)
    assert np.array_equal(perm, np.argsort(key, kind="stable").astype(np.uint64))
    small = np.arange(300, dtype=np.uint64)[::-1].copy()
    perm = np.zeros(300, dtype=np.uint64)
    lib.fastpm_b200_io_argsort_u64(small.ctypes.data, 300, perm.ctypes.data)
    assert np.array_equal(small[perm.astype(np.int64)], np.arange(300, dtype=np.uint64))


def test_write_snapshot_attr_and_string_helpers(lib, run, ref_mod):
    """write_snapshot_attr (io.c:976-998; the CLI stores the parameter file in "Header" with it) on a copy of the reference's
    directory with both libraries: identical attr-v2.  And the string helpers of string.h against the reference's."""
    import shutil
    from oracle import ref
    text = b"nc = 8\nboxsize = 32.0\n-- a parameter file\n\x00"
    frac = C.c_double(0.25)
    outs = []
    for name, L in (("mine", lib), ("ref", ref.lib())):
        top = str(run["tmp"] / ("attr_" + name))
        shutil.copytree(run["ref_dir"], top)
        L.write_snapshot_attr.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_size_t, C.c_int]
        L.write_snapshot_attr(top.encode(), b"Header", b"ParamFile", text, b"S1", len(text), 1)
        L.write_snapshot_attr(top.encode(), b"Header", b"ParticleFraction", C.byref(frac), b"f8", 1, 1)
        outs.append(open(os.path.join(top, "Header", "attr-v2"), "rb").read())
        assert filecmp.cmp(os.path.join(top, "Header", "header"), os.path.join(run["ref_dir"], "Header", "header"), shallow=False)
    assert outs[0] == outs[1] and b"ParamFile" in outs[0] and b"ParticleFraction" in outs[0]
    # fastpm_strsplit: NULL-terminated array of the pieces
    for L in (lib, ref.lib()):
        L.fastpm_strsplit.restype = C.POINTER(C.c_char_p)
        L.fastpm_strsplit.argtypes = [C.c_char_p, C.c_char_p]
    for s, sep in ((b"a,b;;c", b",;"), (b"", b","), (b"nosplit", b","), (b",lead,trail,", b",")):
        got, want = lib.fastpm_strsplit(s, sep), ref.lib().fastpm_strsplit(s, sep)
        i = 0
        while want[i] is not None:
            assert got[i] == want[i]
            i += 1
        assert got[i] is None
    # fastpm_path_ensure_dirname + fastpm_file_get_content
    deep = str(run["tmp"] / "x" / "y" / "z" / "file.txt")
    lib.fastpm_path_ensure_dirname.argtypes = [C.c_char_p]
    lib.fastpm_path_ensure_dirname(deep.encode())
    assert os.path.isdir(os.path.dirname(deep)) and not os.path.exists(deep)
    open(deep, "w").write("0.1\t2.5\n0.2\t1.5\n")
    lib.fastpm_file_get_content.restype = C.c_void_p
    lib.fastpm_file_get_content.argtypes = [C.c_char_p]
    ptr = lib.fastpm_file_get_content(deep.encode())
    assert C.string_at(ptr) == b"0.1\t2.5\n0.2\t1.5\n"
    assert lib.fastpm_file_get_content(b"/nonexistent/file") is None

"""The C-ABI library loads on a CPU-only host, exports every symbol include/icar_b200.h
declares, its pure-host entry points behave like the reference's, and it fails loudly
(never falls back to a CPU path) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import icar_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "icar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icar_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = icar_b200.lib()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libicar_b200.so does not export {n}"
    assert sorted(icar_b200.SYMBOLS) == names


def test_default_options_match_reference_namelist_defaults():
    o = icar_b200.Options()
    # options_obj.f90:1579-1581 (adv_parameters), :1023 (advect_density), :1282-1283 (mp_parameters)
    assert (o.advection, o.microphysics, o.mpdata_order, o.flux_corrected_transport) == (2, 2, 2, 1)
    assert (o.advect_density, o.update_interval, o.top_mp_level) == (0, 0, 0)


def ref_xslab(nx, nz, ny, nimg, img, halo):
    """grid_obj.f90:116-138 (my_n, my_start) + :209-255 (update_with_halos), ximages=nimg, yimages=1."""
    base = nx // nimg
    n = base + (1 if img <= nx % nimg else 0)
    start = (img - 1) * base + min(img - 1, nx % nimg) + 1
    west, east = img == 1, img == nimg
    ims, ime = start, start + n - 1
    ims -= 0 if west else halo
    ime += 0 if east else halo
    its = ims + (1 if west else halo)
    ite = ime - (1 if east else halo)
    return dict(ims=ims, ime=ime, its=its, ite=ite, jms=1, jme=ny, jts=2, jte=ny - 1,
                kms=1, kme=nz, kts=1, kte=nz, ids=1, ide=nx, jds=1, jde=ny, kds=1, kde=nz)


@pytest.mark.parametrize("nx,nimg,halo", [(512, 8, 3), (512, 8, 1), (100, 3, 1), (1024, 4, 3), (37, 5, 2)])
def test_grid_xslab_matches_grid_obj(nx, nimg, halo):
    owned = []
    for img in range(1, nimg + 1):
        g = icar_b200.Grid.xslab(nx, 20, 64, nimg, img, halo)
        for k, v in ref_xslab(nx, 20, 64, nimg, img, halo).items():
            assert getattr(g, k) == v, (img, k)
        assert g.halo_size == halo
        lo = g.ims if img == 1 else g.its
        hi = g.ime if img == nimg else g.ite
        owned.append((lo, hi))
    # the owned ranges tile 1..nx exactly (test_caf_other_image_grids.f90 checks the same)
    assert owned[0][0] == 1 and owned[-1][1] == nx
    for (a, b), (c, d) in zip(owned, owned[1:]):
        assert c == b + 1


def test_grid_yslab_mirrors_xslab():
    for img in (1, 2, 3):
        gx = icar_b200.Grid.xslab(100, 20, 64, 3, img, 3)
        gy = icar_b200.Grid.yslab(64, 20, 100, 3, img, 3)
        assert (gx.ims, gx.ime, gx.its, gx.ite) == (gy.jms, gy.jme, gy.jts, gy.jte)
        assert (gy.ims, gy.ime, gy.its, gy.ite) == (1, 64, 2, 63)
        px, py = icar_b200.halo_plan(gx, img, 3), icar_b200.halo_plan(gy, img, 3)
        assert px["axis"] == "x" and py["axis"] == "y"
        for k in ("send_lo", "send_hi", "recv_lo", "recv_hi", "has_lo", "has_hi", "width"):
            assert px[k] == py[k]
    # exchangeable_obj.f90:303-356: send the h owned columns next to the ghost zone, fill the ghost zone
    g = icar_b200.Grid.xslab(100, 20, 64, 3, 2, 3)
    p = icar_b200.halo_plan(g, 2, 3)
    n = g.nx
    assert (p["send_lo"], p["send_hi"], p["recv_lo"], p["recv_hi"]) == (3, n - 6, 0, n - 3)


def test_grid_xslab_rejects_bad_arguments():
    g = icar_b200.Grid()
    lib = icar_b200.lib()
    assert lib.icar_grid_xslab(10, 5, 10, 0, 1, 1, C.byref(g)) == icar_b200.ERR_INVALID
    assert lib.icar_grid_xslab(10, 5, 10, 4, 5, 1, C.byref(g)) == icar_b200.ERR_INVALID
    assert lib.icar_grid_xslab(3, 5, 10, 4, 1, 1, C.byref(g)) == icar_b200.ERR_INVALID


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a host without a GPU")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(icar_b200.IcarError) as e:
        icar_b200.Domain(icar_b200.Grid.single_image(16, 8, 16))
    assert e.value.code == icar_b200.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The product path must not import, link, include or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "icar_b200")
    pat = re.compile(r"(import|from)\s+oracle|oracle/|oracle\\.|libicar_oracle|_oracle\.hpp|oracle_api")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".f90")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(txt), (dirpath, f)


# ------------------------------------------------------------------ the Fortran shim against the C header
def _fortran_shim():
    return open(os.path.join(ROOT, "icar_b200", "fortran", "icar_b200_mod.f90")).read()


def _c_prototypes():
    """{name: number of parameters} for every function the header declares."""
    src = open(os.path.join(ROOT, "include", "icar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(icar_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        # function-pointer-free header: a plain comma count is the parameter count
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_fortran_shim_binds_existing_symbols_with_matching_arity():
    """No Fortran compiler exists in the image, so the iso_c_binding interfaces are checked textually:
    every `bind(c)` interface must name an exported symbol and take as many dummies as the C prototype."""
    lib = icar_b200.lib()
    protos = _c_prototypes()
    found = re.findall(r"(?:function|subroutine)\s+(icar_[a-z0-9_]+)\s*\(([^)]*)\)\s*bind\(c\)", _fortran_shim(), flags=re.I)
    assert len(found) >= 12
    for name, args in found:
        assert hasattr(lib, name), f"Fortran shim binds {name}, which libicar_b200.so does not export"
        nargs = 0 if not args.strip() else args.count(",") + 1
        assert nargs == protos[name], f"{name}: {nargs} dummies in the shim, {protos[name]} parameters in the header"


def test_fortran_shim_enum_and_struct_layouts_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "icar_b200.h")).read()
    shim = _fortran_shim()
    enum = dict((k, int(v)) for k, v in re.findall(r"(ICAR_VAR_[A-Z_]+)\s*=\s*(\d+)", hdr))
    alias = dict(VAR_TH="ICAR_VAR_POTENTIAL_TEMPERATURE", VAR_QV="ICAR_VAR_WATER_VAPOR", VAR_QC="ICAR_VAR_CLOUD_WATER",
                 VAR_QR="ICAR_VAR_RAIN_IN_AIR", VAR_QS="ICAR_VAR_SNOW_IN_AIR", VAR_PRECIP="ICAR_VAR_PRECIPITATION",
                 VAR_QI="ICAR_VAR_CLOUD_ICE", VAR_NI="ICAR_VAR_ICE_NUMBER", VAR_NR="ICAR_VAR_RAIN_NUMBER",
                 VAR_QG="ICAR_VAR_GRAUPEL_IN_AIR")
    fvars = dict((k, int(v)) for k, v in re.findall(r"\b(VAR_[A-Z_0-9]+)\s*=\s*(\d+)", shim))
    assert len(fvars) >= 20
    for k, v in fvars.items():
        assert enum[alias.get(k, "ICAR_" + k)] == v, k
    # member order of the bind(c) derived types == member order of the C structs
    def c_members(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), re.sub(r"/\*.*?\*/", "", hdr, flags=re.S), flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [n.strip(" *") for n in re.sub(r"^(const\s+)?\w+\s+", "", decl).split(",")]
        return [n.lower() for n in names]

    def f_members(t):
        body = re.search(r"type, bind\(c\) :: %s\n(.*?)end type" % t, shim, flags=re.S | re.I).group(1)
        names = []
        for line in body.splitlines():
            if "::" in line:
                names += [n.strip().lower() for n in line.split("::")[1].split("!")[0].split(",")]
        return names

    for t in ("icar_grid", "icar_options", "icar_mp_options"):
        assert f_members(t) == c_members(t), t


def test_ctypes_structs_match_the_header_member_order():
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "icar_b200.h")).read(), flags=re.S)

    def c_members(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names += [n.strip(" *") for n in re.sub(r"^(const\s+)?\w+\s+", "", decl).split(",")]
        return names

    assert [n for n, _ in icar_b200.Grid._fields_] == c_members("icar_grid")
    assert [n for n, _ in icar_b200.Options._fields_] == c_members("icar_options")
    assert [n for n, _ in icar_b200.MpOptions._fields_] == c_members("icar_mp_options")
    assert [n for n, _ in icar_b200.HostDomain._fields_] == c_members("icar_host_domain")
    assert [n for n, _ in icar_b200.LtOptions._fields_] == c_members("icar_lt_options")
    lt = icar_b200.LtOptions()                         # lt_parameters defaults, options_obj.f90:1448-1479
    assert (lt.variable_N, lt.smooth_nsq, lt.vert_smooth, lt.stability_window_size) == (1, 1, 10, 10)
    assert (lt.n_dir_values, lt.n_nsq_values, lt.n_spd_values, lt.minimum_layer_size) == (24, 5, 6, 100.0)
    assert abs(lt.N_squared - 3e-5) < 1e-12 and abs(lt.linear_update_fraction - 0.2) < 1e-7 and abs(lt.dirmax - 2 * 3.1415927) < 1e-6
    enum = dict((k, int(v)) for k, v in re.findall(r"(ICAR_VAR_[A-Z_]+)\s*=\s*(\d+)", hdr))
    for name, idx in icar_b200.VAR.items():
        assert enum["ICAR_VAR_" + name.upper()] == idx, name


def test_pipeline_band_bounds_cover_the_domain_once():
    """icar_b200.pipeline.band_bounds: owned rows partition 0..ny-1, held rows add the ghost rows inside the domain."""
    from icar_b200.pipeline import band_bounds
    for ny, bands, halo in [(512, 8, 4), (70, 3, 4), (70, 5, 4), (9, 4, 4), (64, 1, 4), (100, 64, 4)]:
        bb = band_bounds(ny, bands, halo)
        owned = [r for (a, b, _, _) in bb for r in range(a, b + 1)]
        assert owned == list(range(ny)), (ny, bands)
        for (a, b, v0, v1) in bb:
            assert v0 == max(0, a - halo) and v1 == min(ny - 1, b + halo)
            assert b - a + 1 >= min(ny, 2 * halo)                  # a band is never thinner than its two halos

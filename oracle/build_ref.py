#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — builds the reference (ilhamv/MC-old) into oracle/_ref/.

The reference's CMake does not configure in this image (cmake 4 rejects
cmake_minimum_required(2.6); HDF5 is absent), so its few source
files are compiled directly with g++ from where they lie under /root/reference.
Nothing from /root/reference is copied into tracked files: outputs (binaries,
a scratch copy of two patched translation units that is deleted after the
compile, and the xs_library data files) go only to oracle/_ref/, which is
git-ignored but travels to the GPU box with gpurun.

Artefacts
  oracle/_ref/MC_ref           unmodified reference sources + h5stub (text output.h5)
  oracle/_ref/MC_ref_patched   same + oracle patches A/B/C so that the
                               slab_analytic / shielding decks run instead of segfaulting
  oracle/_ref/libref_harness.so  reference objects (all but Main/Random)
                               + oracle/ref_harness.cpp: function-level access with injected xi
  oracle/_ref/xs_library/      the reference's cross-section text files (data, read at run time)

Patches (applied to a scratch copy, never to /root/reference):
  A  src/simulator/setup.cpp   a user-defined (capture-only) nuclide gets total=absorb=capture and
                               zero scatter/fission reactions instead of null pointers (F3)
  B  src/simulator/general.cpp guard the fission dispatch when nuclide_nufission() returns null (F4)
  C  src/simulator/setup.cpp   parse xs_library rows per line; a missing 6th column reads as 0 (F5)
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MCB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
STUB = os.path.join(HERE, "h5stub")
CXX = os.environ.get("MCB_CXX", "g++")  # not $CXX: a toolchain wrapper that links libstdc++ statically breaks in-process loading
# x86-64 SSE2 double arithmetic, no FMA contraction: this is what bit-exactness is defined against.
CXXFLAGS = ["-std=c++11", "-O3", "-w", "-fPIC", "-ffp-contract=off", "-I" + STUB, "-I" + os.path.join(REF, "include")]

SRC = ["src/Algorithm.cpp", "src/Distribution.cpp", "src/Entropy.cpp", "src/Estimator.cpp", "src/Geometry.cpp",
       "src/Material.cpp", "src/Nuclide.cpp", "src/Particle.cpp", "src/Random.cpp", "src/Reaction.cpp",
       "src/Source.cpp", "src/XSec.cpp", "src/simulator/fixed_source.cpp", "src/simulator/general.cpp",
       "src/simulator/handler.cpp", "src/simulator/ksearch.cpp", "src/simulator/population_control.cpp",
       "src/simulator/report.cpp", "src/simulator/setup.cpp", "src/simulator/time_dependent.cpp", "Main.cpp"]
# src/simulator/time_dependent.cpp includes <Eigen/Dense> without using it: h5stub/Eigen/Dense is an empty shadow.


def run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def newer(target, *deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def compile_obj(src, obj):
    if not newer(obj, src, os.path.join(STUB, "H5Cpp.h")):
        run([CXX] + CXXFLAGS + ["-c", src, "-o", obj])
    return obj


PATCH_A = r"""
    if( !n.attribute("ZAID") ){ /* oracle patch A */
        if( !n_capture ){ n_capture = std::make_shared<Reaction>( std::make_shared<XSConstant>(0.0) ); }
        n_absorb = n_capture; n_total = n_capture;
        n_scatter = std::make_shared<ReactionScatter>( std::make_shared<XSConstant>(0.0),
                        std::make_shared<DistributionIsotropicScatter>(), n_A );
        std::vector<double> z3(3,0.0), one6(6,1.0), z6(6,0.0);
        n_fission = std::make_shared<ReactionFission>( std::make_shared<XSConstant>(0.0),
                        std::make_shared<XSConstant>(0.0), std::make_shared<DistributionWatt>( z3, z3 ),
                        n_ChiD, std::make_shared<XSConstant>(0.0), one6, z6, z6 );
    }
"""

PATCH_C_OLD = "while ( xs_file >> c1 >> c2 >> c3 >> c4 >> c5 >> c6 ){"
PATCH_C_NEW = ("std::string oracle_line; std::getline( xs_file, oracle_line ); /* oracle patch C */\n"
               "        while ( std::getline( xs_file, oracle_line ) ){\n"
               "            std::istringstream oracle_ls( oracle_line );\n"
               "            if( !( oracle_ls >> c1 >> c2 >> c3 >> c4 >> c5 ) ){ continue; }\n"
               "            if( !( oracle_ls >> c6 ) ){ c6 = 0.0; }")


def patched_sources(scratch):
    os.makedirs(scratch, exist_ok=True)
    setup = open(os.path.join(REF, "src/simulator/setup.cpp")).read()
    assert PATCH_C_OLD in setup
    setup = setup.replace(PATCH_C_OLD, PATCH_C_NEW)
    anchor = "    N = std::make_shared<Nuclide> ( n_name, n_A, n_capture, n_scatter,"
    assert anchor in setup
    setup = setup.replace(anchor, PATCH_A + anchor)
    setup = setup.replace('#include <fstream> ', '#include <fstream>\n#include <sstream>')
    general = open(os.path.join(REF, "src/simulator/general.cpp")).read()
    old_b = "    if( ksearch ){ \n        implicit_fission_ksearch( P, bank_nu, N_fission );"
    assert old_b in general
    general = general.replace(old_b, "    if( !N_fission ){ /* oracle patch B */ }\n    else if( ksearch ){ \n"
                                     "        implicit_fission_ksearch( P, bank_nu, N_fission );")
    p_setup = os.path.join(scratch, "setup_patched.cpp")
    p_general = os.path.join(scratch, "general_patched.cpp")
    open(p_setup, "w").write(setup)
    open(p_general, "w").write(general)
    return p_setup, p_general


def build(force=False):
    if not os.path.isdir(REF):
        print("[oracle/build_ref] %s not present: keeping prebuilt oracle/_ref as is" % REF)
        return False
    os.makedirs(OBJ, exist_ok=True)
    # data files the decks read at run time (./xs_library relative to CWD)
    xs_dst = os.path.join(OUT, "xs_library")
    if not os.path.isdir(xs_dst):
        shutil.copytree(os.path.join(REF, "xs_library"), xs_dst)

    targets = [os.path.join(OUT, t) for t in ("MC_ref", "MC_ref_patched", "libref_harness.so")]
    deps = [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "build_ref.py"), os.path.join(STUB, "H5Cpp.h"),
            os.path.join(STUB, "ref_stubs.cpp")]
    if not force and all(newer(t, *deps) for t in targets):
        print("[oracle/build_ref] oracle/_ref is up to date")
        return True

    pugi = os.path.join(OBJ, "pugixml.o")
    if not os.path.exists(pugi):
        run([CXX, "-std=c++11", "-O2", "-w", "-fPIC", "-I" + os.path.join(REF, "include"), "-c",
             os.path.join(REF, "src/pugixml/pugixml.cpp"), "-o", pugi])
    objs = {}
    for s in SRC:
        objs[s] = compile_obj(os.path.join(REF, s), os.path.join(OBJ, s.replace("/", "_")[:-4] + ".o"))
    stubs = compile_obj(os.path.join(STUB, "ref_stubs.cpp"), os.path.join(OBJ, "ref_stubs.o"))

    run([CXX, "-O3"] + list(objs.values()) + [stubs, pugi, "-o", targets[0]])

    scratch = os.path.join(OBJ, "scratch")
    p_setup, p_general = patched_sources(scratch)
    o_setup = os.path.join(OBJ, "setup_patched.o")
    o_general = os.path.join(OBJ, "general_patched.o")
    run([CXX] + CXXFLAGS + ["-c", p_setup, "-o", o_setup])
    run([CXX] + CXXFLAGS + ["-c", p_general, "-o", o_general])
    shutil.rmtree(scratch)
    pobjs = dict(objs)
    pobjs["src/simulator/setup.cpp"] = o_setup
    pobjs["src/simulator/general.cpp"] = o_general
    run([CXX, "-O3"] + list(pobjs.values()) + [stubs, pugi, "-o", targets[1]])

    # function-level harness: patched setup/general (so every deck loads), our injectable Urand instead of Random.o
    hobjs = [o for s, o in pobjs.items() if s not in ("Main.cpp", "src/Random.cpp")]
    harness = os.path.join(OBJ, "ref_harness.o")
    run([CXX] + CXXFLAGS + ["-c", os.path.join(HERE, "ref_harness.cpp"), "-o", harness])
    run([CXX, "-shared", "-O3"] + hobjs + [harness, stubs, pugi, "-o", targets[2]])
    print("[oracle/build_ref] built", ", ".join(os.path.basename(t) for t in targets))
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)

#!/bin/bash
# Builds the unmodified reference out of tree and installs it into baseline/_ref (git-ignored,
# shipped to the GPU box by gpurun).  TEST / BASELINE INFRASTRUCTURE: the product never imports it.
#   * sources stay under /root/reference; a scratch copy is compiled in $TMP (the reference's
#     setup.py writes into its tree)
#   * `pip install` is not used: one optional extension (lib/_transformations) does not compile
#     against numpy 2.3 and aborts the build; the other 13 are built by build_ext and copied in
#   * baseline/_ref receives the package (.py + built .so, no C / Cython sources) and four stub
#     modules for imports the path never exercises (mmtf, fasteners, mda_xdrlib, matplotlib)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFROOT:-/root/reference}"
TMP="${MDA_REF_TMP:-/tmp/mda_ref}"
[ -d "$REF/package/MDAnalysis" ] || { echo "reference checkout absent: keeping baseline/_ref as it is"; exit 0; }
if [ ! -d "$TMP/package/build" ]; then
  mkdir -p "$TMP" && rm -rf "$TMP/package" && cp -r "$REF/package" "$TMP/"
  (cd "$TMP/package" && CC=/usr/bin/gcc CXX=/usr/bin/g++ python setup.py build_ext -j8 > "$TMP/build.log" 2>&1 || true)
fi
LIB="$(ls -d "$TMP"/package/build/lib.* | head -1)"
[ -n "$(find "$LIB" -name 'c_distances*.so' | head -1)" ] || { echo "reference build failed, see $TMP/build.log"; exit 1; }
rm -rf "$HERE/_ref" && mkdir -p "$HERE/_ref"
cp -r "$TMP/package/MDAnalysis" "$HERE/_ref/"
(cd "$LIB" && find MDAnalysis -name '*.so' -exec cp {} "$HERE/_ref/{}" \;)
find "$HERE/_ref/MDAnalysis" \( -name '*.c' -o -name '*.cpp' -o -name '*.pyx' -o -name '*.pxd' -o -name '*.h' -o -name '__pycache__' \) -prune -exec rm -rf {} +
mkdir -p "$HERE/_ref/mda_xdrlib" "$HERE/_ref/matplotlib"
printf 'class MMTFDecoder: pass\ndef parse(*a, **k): raise NotImplementedError\ndef fetch(*a, **k): raise NotImplementedError\n' > "$HERE/_ref/mmtf.py"
printf 'class InterProcessLock:\n    def __init__(self, *a, **k): pass\n    def __enter__(self): return self\n    def __exit__(self, *a): return False\n' > "$HERE/_ref/fasteners.py"
: > "$HERE/_ref/mda_xdrlib/__init__.py"
printf 'class Unpacker:\n    def __init__(self, *a, **k): raise NotImplementedError\nclass Packer:\n    def __init__(self, *a, **k): raise NotImplementedError\n' > "$HERE/_ref/mda_xdrlib/xdrlib.py"
: > "$HERE/_ref/matplotlib/__init__.py"
: > "$HERE/_ref/matplotlib/pyplot.py"
# the reference's own tests for the hot path: the MDAnalysisTests package (python files) and the
# ~19 MB of data files those test modules load; used by tests/test_gpu_reference_testsuite.py
TS="$REF/testsuite/MDAnalysisTests"
if [ -d "$TS" ]; then
  mkdir -p "$HERE/_ref/MDAnalysisTests/data"
  (cd "$TS" && find . -name '*.py' -not -path './data/*' | while read f; do mkdir -p "$HERE/_ref/MDAnalysisTests/$(dirname "$f")"; cp "$f" "$HERE/_ref/MDAnalysisTests/$f"; done)
  [ -f "$TS/data/__init__.py" ] && cp "$TS/data/__init__.py" "$HERE/_ref/MDAnalysisTests/data/"
  for f in analysis/align_bound.pdb.gz analysis/align_unbound.pdb.gz adk_open.crd adk_dims.dcd test.fasta adk_oplsaa.gro \
           analysis/YiiP_lipids.gro.gz martini_dppc_chol_bilayer.gro adk_oplsaa.pdb A6PA6_alpha.pdb cobrotoxin.pdb adk.psf \
           surface.pdb.bz2 surface.trr adk_oplsaa.tpr adk_oplsaa.xtc analysis/YiiP_lipids.xtc contacts/2F4K_qlist5_remap.dat \
           contacts/villin_folded.gro.bz2 contacts/villin_unfolded.gro.bz2 adk_oplsaa_CA_rmsf.npy two_water_gro.gro \
           tip125_tric_C36.psf tip125_tric_C36.dcd dppc_vesicle_hg.gro adk_open.pdb trz_psf.psf trzfile.trz 1osm.pdb.gz \
           watdyn.psf watdyn.dcd namd_cgenff.psf namd_cgenff.pdb lammps/a_lot_of_bond_types.data 19hc.pdb.gz 1a28.pdb.gz adk_oplsaa_GLY_rama.npy \
           adk_oplsaa_LYS_janin.npy adk_oplsaa_dihedral.npy adk_oplsaa_dihedral_list.npy adk_oplsaa_janin.npy adk_oplsaa_rama.npy; do
    mkdir -p "$HERE/_ref/MDAnalysisTests/data/$(dirname "$f")"
    cp "$TS/data/$f" "$HERE/_ref/MDAnalysisTests/data/$f"
  done
fi
PYTHONPATH="$HERE/_ref" python -W ignore -c "import MDAnalysis, MDAnalysis.lib.distances as d; print('baseline/_ref: MDAnalysis', MDAnalysis.__version__, list(d._distances))"

"""Deck writers for tests and the bench: they emit input.xml text in the reference's deck grammar
 for the physics of BASELINE.json's configs, with the history counts as parameters.
The physics values (densities, radii, source) are those of the reference's example decks
(examples/<name>/input.xml, cited per function) so results can be compared with its documented numbers.
"""
import os

HEAD = "<?xml version = '1.0' encoding = 'UTF-8'?>\n\n"


def slab(samples=100000):
    """examples/slab_analytic/input.xml: two purely absorbing slabs, leak through x=5 = exp(-4.2)."""
    return HEAD + f"""
<simulation>
    <description name="Simple Slabs" samples="{samples:g}"/>
</simulation>
<nuclides>
    <nuclide name="nuc1"> <capture xs="1.0"/> </nuclide>
    <nuclide name="nuc2"> <capture xs="0.5"/> </nuclide>
    <nuclide name="nuc3"> <capture xs="2.0"/> </nuclide>
</nuclides>
<materials>
    <material name="mat1">
        <nuclide name="nuc1" density="0.5"/>
        <nuclide name="nuc2" density="1.0"/>
        <nuclide name="nuc3" density="0.1"/>
    </material>
    <material name="mat2">
        <nuclide name="nuc1" density="0.5"/>
        <nuclide name="nuc2" density="0.5"/>
    </material>
</materials>
<surfaces>
    <plane_x name="px1" x="0.0"/>
    <plane_x name="px2" x="1.0"/>
    <plane_x name="px3" x="5.0"/>
</surfaces>
<cells>
    <cell name="slab1" material="mat1">
        <surface name="px1" sense="+1"/>
        <surface name="px2" sense="-1"/>
    </cell>
    <cell name="slab2" material="mat2">
        <surface name="px2" sense="+1"/>
        <surface name="px3" sense="-1"/>
    </cell>
    <cell name="left outside" importance="0.0">
        <surface name="px1" sense="-1"/>
    </cell>
    <cell name="right outside" importance="0.0">
        <surface name="px3" sense="+1"/>
    </cell>
</cells>
<distributions>
    <delta name="dir" datatype="point" x = "1.0" y = "0.0" z = "0.0"/>
</distributions>
<sources>
    <point x="1e-9" y="0.0" z="0.0" direction="dir"/>
</sources>
<estimators>
    <estimator name="leak_rate" scores="cross">
        <surface name="px3"/>
    </estimator>
</estimators>
"""


def heu_sphere(samples=1000, active=5, passive=3, source_tag="point", entropy=False, estimators=False):
    """examples/HEU_sphere_criticality/input.xml: bare HEU sphere r=7.68 cm, k-eigenvalue.
    source_tag="source" writes the deck's own <source position=...> form (rejected by the reference, F6)."""
    src = ('<point x="0.0" y="0.0" z="0.0" direction="dir" energy="enrg"/>' if source_tag == "point"
           else '<source position="pos" direction="dir" energy="enrg"/>')
    ent = """
    <entropy>
        <x min="-7.68" max="7.68" step="8"/>
        <y min="-7.68" max="7.68" step="8"/>
        <z min="-7.68" max="7.68" step="8"/>
    </entropy>""" if entropy else ""
    est = """
<estimators>
    <estimator name="sphere_rates" scores="flux fission nu-fission absorption">
        <cell name="sphere"/>
        <filter type="energy" grid="1e-5 1.0 1e3 1e5 1e6 2e6 5e6 2e7"/>
    </estimator>
    <estimator name="sphere_coll" type="C" scores="flux total">
        <cell name="sphere"/>
    </estimator>
    <estimator name="leak" scores="cross">
        <surface name="suspicious_sphere"/>
    </estimator>
</estimators>""" if estimators else ""
    return HEAD + f"""
<simulation>
    <description name="HEU Sphere" samples="{samples:g}"/>
    <ksearch active_cycles="{active}" passive_cycles="{passive}"/>{ent}
</simulation>
<distributions>
    <delta name="pos" datatype="point" x="0.0" y="0.0" z="0.0" />
    <isotropic name="dir" datatype="point" />
    <delta name="enrg" datatype="double" val="14.0e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
    <nuclide name="U-238" ZAID="092238"/>
</nuclides>
<materials>
    <material name="HEU">
        <nuclide name="U-235" density="0.0455112"/>
        <nuclide name="U-238" density="0.0033823"/>
    </material>
</materials>
<surfaces>
    <sphere name="suspicious_sphere"  x="0.0" y="0.0" z="0.0" r="7.68"/>
</surfaces>
<cells>
    <cell name="sphere" material="HEU">
        <surface name="suspicious_sphere" sense="-1" />
    </cell>
    <cell name="graveyard" importance="0.0">
        <surface name="suspicious_sphere" sense="+1" />
    </cell>
</cells>
<sources>
    {src}
</sources>{est}
"""


def ucube(samples=1000, active=5, passive=3):
    """examples/UCube/input.xml: U-235 box with reflective/vacuum planes and the only <entropy> mesh."""
    return HEAD + f"""
<simulation>
    <description name="Infinite U235 Cube" samples="{samples:g}"/>
    <ksearch active_cycles="{active}" passive_cycles="{passive}"/>
    <entropy>
        <x min="0.0" max="12.0" step="12"/>
        <y min="0.0" max="5.0" step="5"/>
        <z min="0.0" max="6.0" step="6"/>
    </entropy>
</simulation>
<distributions>
    <isotropic name="dir" datatype="point" />
    <delta name="enrg" datatype="double" val="14.0e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
</nuclides>
<materials>
    <material name="Fuel">
        <nuclide name="U-235" density="0.0508"/>
    </material>
</materials>
<surfaces>
    <plane_x name="px1" x="12.0" bc="vacuum"/>
    <plane_x name="px2" x="-0.0" bc="reflective"/>
    <plane_y name="py1" y="5.0" bc="vacuum"/>
    <plane_y name="py2" y="-0.0" bc="reflective"/>
    <plane_z name="pz1" z="6.0" bc="vacuum"/>
    <plane_z name="pz2" z="-0.0" bc="reflective"/>
</surfaces>
<cells>
    <cell name="UCube" material="Fuel">
        <surface name="px1" sense="-1" />
        <surface name="px2" sense="+1" />
        <surface name="py1" sense="-1" />
        <surface name="py2" sense="+1" />
        <surface name="pz1" sense="-1" />
        <surface name="pz2" sense="+1" />
    </cell>
</cells>
<sources>
    <point x="6.0" y="2.5" z="3.0" direction="dir" energy="enrg"/>
</sources>
"""


def gcr(samples=400, active=5, passive=2, trmm=False, groups=20):
    """examples/infinite_GCR_TRMM/input.xml: infinite graphite-moderated medium (4 nuclides), reflective planes.
    groups=100 is examples/infinite_GCR_TRMM_100 (8 * 100^2 + 15 * 100 = 81500 tallies)."""
    t = f"""
<trmm>
    <cell name="infinity"/>
    <filter type="energy" grid_lethargy="1E-5 2E7 {groups}"/>
</trmm>""" if trmm else ""
    return HEAD + f"""
<simulation>
    <description name="Infinite GCR" samples="{samples:g}"/>
    <ksearch active_cycles="{active}" passive_cycles="{passive}"/>
</simulation>
<distributions>
    <isotropic name="dir" datatype="point" />
    <delta name="enrg" datatype="double" val="14.0e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
    <nuclide name="U-238" ZAID="092238"/>
    <nuclide name="O-16"  ZAID="008016"/>
    <nuclide name="C"     ZAID="006000"/>
</nuclides>
<materials>
    <material name="Fuel">
        <nuclide name="U-235" density="0.0000402"/>
        <nuclide name="U-238" density="0.0009061"/>
        <nuclide name="O-16"  density="0.0018927"/>
        <nuclide name="C"     density="0.0757080"/>
    </material>
</materials>
<surfaces>
    <plane_x name="px1" x="100.0"  bc="reflective"/>
    <plane_x name="px2" x="-100.0" bc="reflective"/>
</surfaces>
<cells>
    <cell name="infinity" material="Fuel">
        <surface name="px1" sense="-1" />
        <surface name="px2" sense="+1" />
    </cell>
</cells>
<sources>
    <point x="0.0" y="0.0" z="0.0" direction="dir" energy="enrg"/>
</sources>{t}
"""


def gcr_td(samples=2500, times="3e-8 15e-8 4e-6 100e-6", linear=None, comb=None, groups=50):
    """examples/infinite_GCR_TD/input.xml: the same medium in time-dependent mode (<tdmc>): a 14.1 MeV pulse at t = 0,
    the spectrum scored at the census times through the <tdmc/> filter.  linear = "a n b" uses time_linear instead;
    comb = (bank_max, teeth) adds the particle comb (examples/infinite_GCR_TD_sub)."""
    grid = f'time_linear="{linear}"' if linear else f'time="{times}"'
    ctrl = "" if comb is None else f"""
<population_control>
    <particle_comb bank_max="{comb[0]}" teeth="{comb[1]}"/>
</population_control>"""
    return HEAD + f"""
<simulation>
    <description name="Infinite GCR" samples="{samples:g}"/>
    <tdmc {grid}/>
</simulation>{ctrl}
<distributions>
    <isotropic name="dir" datatype="point" />
    <delta name="enrg" datatype="double" val="14.1e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
    <nuclide name="U-238" ZAID="092238"/>
    <nuclide name="O-16"  ZAID="008016"/>
    <nuclide name="C"     ZAID="006000"/>
</nuclides>
<materials>
    <material name="Fuel">
        <nuclide name="U-235" density="0.0000402"/>
        <nuclide name="U-238" density="0.0009061"/>
        <nuclide name="O-16"  density="0.0018927"/>
        <nuclide name="C"     density="0.0757080"/>
    </material>
</materials>
<surfaces>
    <plane_x name="px1" x="100.0"  bc="reflective"/>
    <plane_x name="px2" x="-100.0" bc="reflective"/>
</surfaces>
<cells>
    <cell name="infinity" material="Fuel">
        <surface name="px1" sense="-1" />
        <surface name="px2" sense="+1" />
    </cell>
</cells>
<sources>
    <point x="0.0" y="0.0" z="0.0" direction="dir" energy="enrg"/>
</sources>
<estimators>
    <estimator name="spectrum" scores="flux">
        <cell name="infinity"/>
        <filter type="energy" grid_lethargy="1E-5 2E7 {groups}"/>
        <tdmc/>
    </estimator>
    <estimator name="rates" type="C" scores="flux fission">
        <cell name="infinity"/>
    </estimator>
</estimators>
"""


def shielding(samples=20000, split=False):
    """examples/shielding_vReduction/input.xml: water/B4C shield with a He-3 detector (6 nuclides, 10 cells).
    split=True raises the importance of the cells towards the detector so that splitting is exercised."""
    i3, i4, idet = ("2.0", "4.0", "4.0") if split else ("1.0", "1.0", "1.0")
    return HEAD + f"""
<simulation>
    <description name="Shielding w/ Variance Reduction - 1" samples="{samples:g}"/>
</simulation>
<nuclides>
    <nuclide name="O16" ZAID="008016"/>
    <nuclide name="H1"  ZAID="001001"/>
    <nuclide name="He3" ZAID="002003"/>
    <nuclide name="C0"  ZAID="006000"/>
    <nuclide name="B10" ZAID="005010"/>
    <nuclide name="B11" ZAID="005011"/>
</nuclides>
<materials>
    <material name="B4C">
        <nuclide name="B10" density="0.0219716"/>
        <nuclide name="B11" density="0.0878864"/>
        <nuclide name="C0"  density="0.027468"/>
    </material>
    <material name="H2O">
        <nuclide name="H1"  density="0.066733"/>
        <nuclide name="O16" density="0.033368"/>
    </material>
    <material name="mat_detector">
        <nuclide name="He3" density="0.00002501"/>
    </material>
</materials>
<surfaces>
    <plane_x    name="px1" x="0.0"/>
    <plane_x    name="px2" x="4.0"/>
    <plane_x    name="px3" x="5.0"/>
    <plane_x    name="px4" x="9.0"/>
    <plane_y    name="py1" y="0.0"/>
    <plane_y    name="py2" y="3.0"/>
    <plane_y    name="py3" y="6.0"/>
    <cylinder_z name="cz1" x="6.5" y="1.5" r="0.5"/>
</surfaces>
<cells>
    <cell name="water1" material="H2O" importance="1.0">
        <surface name="px1" sense="+1"/> <surface name="px2" sense="-1"/>
        <surface name="py1" sense="+1"/> <surface name="py2" sense="-1"/>
    </cell>
    <cell name="water2" material="H2O" importance="1.0">
        <surface name="px1" sense="+1"/> <surface name="px2" sense="-1"/>
        <surface name="py2" sense="+1"/> <surface name="py3" sense="-1"/>
    </cell>
    <cell name="water3" material="H2O" importance="{i3}">
        <surface name="px2" sense="+1"/> <surface name="px4" sense="-1"/>
        <surface name="py2" sense="+1"/> <surface name="py3" sense="-1"/>
    </cell>
    <cell name="water4" material="H2O" importance="{i4}">
        <surface name="px3" sense="+1"/> <surface name="px4" sense="-1"/>
        <surface name="py1" sense="+1"/> <surface name="py2" sense="-1"/>
        <surface name="cz1" sense="+1"/>
    </cell>
    <cell name="shield" material="B4C" importance="{i3}">
        <surface name="px2" sense="+1"/> <surface name="px3" sense="-1"/>
        <surface name="py1" sense="+1"/> <surface name="py2" sense="-1"/>
    </cell>
    <cell name="detector" material="mat_detector" importance="{idet}">
        <surface name="cz1" sense="-1"/>
    </cell>
    <cell name="left outside" importance="0.0"> <surface name="px1" sense="-1"/> </cell>
    <cell name="right outside" importance="0.0"> <surface name="px4" sense="+1"/> </cell>
    <cell name="down outside" importance="0.0"> <surface name="py1" sense="-1"/> </cell>
    <cell name="up outside" importance="0.0"> <surface name="py3" sense="+1"/> </cell>
</cells>
<estimators>
    <estimator name="detector_response" scores="flux absorption">
        <cell  name="detector"/>
    </estimator>
</estimators>
<distributions>
    <delta name="enrg" datatype="double" val="2.0e6"/>
</distributions>
<sources>
    <point x="1.5"  y="1.5"  z="0.0" energy="enrg"/>
</sources>
"""


def fixed_source_fissile(samples=2000, comb=None):
    """A fixed-source deck in fissile material (same-history fission secondaries, fixed_source.cpp:12-22),
    generic plane + cylinder_x surfaces, a uniform source energy and an independentXYZ direction, with
    TL / collision / surface estimators and an energy filter — exercises the paths the example decks leave out.
    comb = (bank_max, teeth) adds <population_control><particle_comb/> (population_control.cpp:55-84)."""
    ctrl = "" if comb is None else f"""
<population_control>
    <particle_comb bank_max="{comb[0]}" teeth="{comb[1]}"/>
</population_control>"""
    return HEAD + f"""
<simulation>
    <description name="subcritical block" samples="{samples:g}"/>
</simulation>{ctrl}
<distributions>
    <uniform name="ux" datatype="double" a="0.2" b="0.9"/>
    <uniform name="uy" datatype="double" a="-0.3" b="0.3"/>
    <delta   name="uz" datatype="double" val="0.1"/>
    <independentXYZ name="dir" datatype="point" x="ux" y="uy" z="uz"/>
    <uniform name="enrg" datatype="double" a="1.0e5" b="3.0e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
    <nuclide name="C"     ZAID="006000"/>
</nuclides>
<materials>
    <material name="fuel">
        <nuclide name="U-235" density="0.02"/>
        <nuclide name="C"     density="0.04"/>
    </material>
    <material name="graphite">
        <nuclide name="C" density="0.08"/>
    </material>
</materials>
<surfaces>
    <plane name="p1" a="1.0" b="0.0" c="0.0" d="-4.0" bc="vacuum"/>
    <plane name="p2" a="1.0" b="0.2" c="0.0" d="5.0" bc="vacuum"/>
    <cylinder_x name="cx" y="0.0" z="0.0" r="3.0"/>
    <cylinder_x name="cxo" y="0.0" z="0.0" r="6.0" bc="vacuum"/>
</surfaces>
<cells>
    <cell name="core" material="fuel">
        <surface name="p1" sense="+1"/> <surface name="p2" sense="-1"/> <surface name="cx" sense="-1"/>
    </cell>
    <cell name="refl" material="graphite">
        <surface name="p1" sense="+1"/> <surface name="p2" sense="-1"/>
        <surface name="cx" sense="+1"/> <surface name="cxo" sense="-1"/>
    </cell>
</cells>
<sources>
    <point x="0.0" y="0.0" z="0.0" direction="dir" energy="enrg"/>
</sources>
<estimators>
    <estimator name="core_tl" scores="flux fission nu-fission capture scatter total absorption">
        <cell name="core"/> <cell name="refl"/>
        <filter type="energy" grid_lethargy="1e-3 2e7 6"/>
    </estimator>
    <estimator name="core_c" type="C" scores="flux absorption">
        <cell name="core"/>
    </estimator>
    <estimator name="interface" scores="cross flux">
        <surface name="cx"/> <surface name="p2"/>
    </estimator>
</estimators>
"""


def heu_leakage(samples=2000):
    """examples/HEU_sphere_leakage/input.xml: fixed 14 MeV point source in the bare HEU sphere, time-binned leakage
    (surface estimator with a time filter, Estimator.cpp:199-246); plus a track-length estimator with time x energy
    filters (tracks split over several time bins) and a collision estimator with a time filter."""
    return HEAD + f"""
<simulation>
    <description name="Fission Sphere" samples="{samples:g}"/>
</simulation>
<distributions>
    <isotropic name="dir" datatype="point" />
    <delta name="enrg" datatype="double" val="14.0e6"/>
</distributions>
<nuclides>
    <nuclide name="U-235" ZAID="092235"/>
    <nuclide name="U-238" ZAID="092238"/>
</nuclides>
<materials>
    <material name="HEU">
        <nuclide name="U-235" density="0.0455112"/>
        <nuclide name="U-238" density="0.0033823"/>
    </material>
</materials>
<surfaces>
    <sphere name="suspicious_sphere"  x="0.0" y="0.0" z="0.0" r="7.68"/>
</surfaces>
<cells>
    <cell name="sphere" material="HEU">
        <surface name="suspicious_sphere" sense="-1" />
    </cell>
    <cell name="graveyard" importance="0.0">
        <surface name="suspicious_sphere" sense="+1" />
    </cell>
</cells>
<estimators>
    <estimator name="sphere_leakage" scores="cross">
        <surface name="suspicious_sphere"/>
        <filter type="time" grid_linear="0.0 1e-10 1e-8"/>
    </estimator>
    <estimator name="sphere_flux_t" scores="flux fission">
        <cell name="sphere"/>
        <filter type="time" grid_linear="0.0 2e-10 6e-9"/>
        <filter type="energy" grid="1e-5 1e5 1e6 5e6 2e7"/>
    </estimator>
    <estimator name="sphere_coll_t" type="C" scores="flux">
        <cell name="sphere"/>
        <filter type="time" grid="0.0 1e-9 3e-9 1e-8"/>
    </estimator>
</estimators>
<sources>
    <point x="0.0" y="0.0" z="0.0" direction="dir" energy="enrg"/>
</sources>
"""


def write(dirpath, text):
    os.makedirs(dirpath, exist_ok=True)
    with open(os.path.join(dirpath, "input.xml"), "w") as f:
        f.write(text)
    return dirpath


def sphere_detection(samples=100000):
    """examples/sphere_detection/input.xml: a graphite sphere lit by a mono-directional 1 MeV disk source, seen by a He-3
    tube inside a polyethylene moderator.  The deck of the reference's MCNP6 integral test
    (test/test_integral_Simulator.cpp:10-19: detector_TL absorption = 6.9276e-5 per source particle); its <disk_z> source is
    rejected by the reference's own loader (setup.cpp:1051-1063) and sampled here as include/mcb200.h says."""
    return HEAD + f"""
<simulation>
    <description name="Detecting a Sphere" samples="{samples:g}"/>
</simulation>
<nuclides>
    <nuclide name="C0"  ZAID="006000"/>
    <nuclide name="H1"  ZAID="001001"/>
    <nuclide name="He3" ZAID="002003"/>
</nuclides>
<materials>
    <material name="polyethylene">
        <nuclide name="C0" density="0.039929"/>
        <nuclide name="H1" density="0.079855"/>
    </material>
    <material name="helium3">
        <nuclide name="He3" density="0.00002501"/>
    </material>
    <material name="graphite">
        <nuclide name="C0" density="0.100280"/>
    </material>
</materials>
<surfaces>
    <sphere     name="sp1"  x="0.0" y="0.0" z="0.0" r="4.0"/>
    <plane_x    name="px1"  x="9.0"/>
    <plane_x    name="px2"  x="24.0"/>
    <cylinder_x name="cx1"  y="0.0" z="0.0" r="5.5"/>
    <plane_x    name="px11" x="14.0"/>
    <plane_x    name="px22" x="19.0"/>
    <cylinder_x name="cx11" y="0.0" z="0.0" r="0.5"/>
</surfaces>
<cells>
    <cell name="sphere" material="graphite">
        <surface name="sp1" sense="-1"/>
    </cell>
    <cell name="detector" material="helium3">
        <surface name="px11" sense="+1"/>
        <surface name="px22" sense="-1"/>
        <surface name="cx11" sense="-1"/>
    </cell>
    <cell name="moderator left" material="polyethylene">
        <surface name="px1"  sense="+1"/>
        <surface name="px11" sense="-1"/>
        <surface name="cx1"  sense="-1"/>
    </cell>
    <cell name="moderator mid" material="polyethylene">
        <surface name="px11" sense="+1"/>
        <surface name="px22" sense="-1"/>
        <surface name="cx1" sense="-1"/>
        <surface name="cx11" sense="+1"/>
    </cell>
    <cell name="moderator right" material="polyethylene">
        <surface name="px22" sense="+1"/>
        <surface name="px2"  sense="-1"/>
        <surface name="cx1"  sense="-1"/>
    </cell>
    <cell name="left vacuum">
        <surface name="sp1" sense="+1"/>
        <surface name="px1" sense="-1"/>
    </cell>
    <cell name="middle vacuum" importance="0.0">
        <surface name="cx1" sense="+1"/>
        <surface name="px1" sense="+1"/>
        <surface name="px2" sense="-1"/>
    </cell>
    <cell name="right vacuum" importance="0.0">
        <surface name="px2" sense="+1"/>
    </cell>
</cells>
<estimators>
    <estimator name="detector_TL" scores="flux absorption" type="TL">
        <cell name="detector"/>
    </estimator>
</estimators>
<estimators>
    <estimator name="detector_C" scores="flux absorption" type="C">
        <cell name="detector"/>
    </estimator>
</estimators>
<distributions>
    <delta name="dir" datatype="point" x = "0.0" y = "0.0" z = "1.0"/>
    <delta name="enrg" datatype="double" val="1.0e6"/>
</distributions>
<sources>
    <disk_z x="-1.0"  y="0.0" z="-5.0" r="2.0" direction="dir" energy="enrg"/>
</sources>
"""


def slab_overlap(samples=20000):
    """The slab of slab_analytic with a void cell that OVERLAPS the one-surface outside cell and comes before it in deck
    order: a particle leaving through x = 5 is found by search_cell (general.cpp:26-34, first match in deck order) in
    "catch", flies on to y = 1000 and only there enters "right outside".  A crossing shortcut that took the one-surface
    cell behind x = 5 for granted would end the history at x = 5 and the second estimator would stay empty."""
    x = slab(samples)
    for old, new in [('<plane_x name="px3" x="5.0"/>', '<plane_x name="px3" x="5.0"/>\n    <plane_y name="py" y="1000.0"/>'),
                     ('<cell name="left outside" importance="0.0">',
                      '<cell name="catch">\n        <surface name="py" sense="-1"/>\n    </cell>\n    <cell name="left outside" importance="0.0">'),
                     ('x = "1.0" y = "0.0" z = "0.0"', 'x = "0.8" y = "0.6" z = "0.0"'),
                     ('</estimators>', '    <estimator name="far_plane" scores="cross">\n        <surface name="py"/>\n    </estimator>\n</estimators>')]:
        assert old in x, old
        x = x.replace(old, new)
    return x

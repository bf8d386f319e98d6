"""The command-line keywords of SOS_ABS_MAIN -- the 98 names its argument loop compares with (SOS_ABS_MAIN.F:1499-2200; documented at
:213-912, where -AER.SF.RH appears as "-AER.SF.HR") -- with their value types, the defaults of inc/SOS.h, and a parser
that accepts exactly the reference's `-KEYWORD value` pairs.  The keyword set is the drop-in surface; which combinations the
device front end can run is decided in frontend.py, not here."""

F, I, S = "f", "i", "s"
KEYWORDS = {
    "-SOS_Main.Wa": F, "-SOS_Main.ResRoot": S, "-SOS_Main.Log": S,
    "-ANG.Rad.NbGauss": I, "-ANG.Rad.UserAngFile": S, "-ANG.Thetas": F, "-ANG.Rad.ResFile": S, "-ANG.Aer.NbGauss": I,
    "-ANG.Aer.UserAngFile": S, "-ANG.Aer.ResFile": S, "-ANG.Log": S,
    "-SURF.Dir": S, "-SURF.Type": I, "-SURF.Log": S, "-SURF.Ind": F, "-SURF.Glitter.Wind": F, "-SURF.Roujean.K0": F,
    "-SURF.Roujean.K1": F, "-SURF.Roujean.K2": F, "-SURF.Nadal.Alpha": F, "-SURF.Nadal.Beta": F, "-SURF.Maignan.C": F,
    "-SURF.Alb": F, "-SURF.File": S,
    "-AER.Waref": F, "-AER.AOTref": F, "-AER.Tronca": I, "-AER.Log": S, "-AER.MieLog": S, "-AER.DirMie": S, "-AER.ResFile": S,
    "-AER.UserFile": S, "-AER.Model": I,
    "-AER.MMD.MRwa": F, "-AER.MMD.MIwa": F, "-AER.MMD.MRwaref": F, "-AER.MMD.MIwaref": F, "-AER.MMD.SDtype": I,
    "-AER.MMD.LNDradius": F, "-AER.MMD.LNDvar": F, "-AER.MMD.JD.slope": F, "-AER.MMD.JD.rmin": F, "-AER.MMD.JD.rmax": F,
    "-AER.WMO.Model": I, "-AER.WMO.DL": F, "-AER.WMO.WS": F, "-AER.WMO.OC": F, "-AER.WMO.SO": F,
    "-AER.SF.Model": I, "-AER.SF.RH": F,
    "-AER.BMD.VCdef": I, "-AER.BMD.CoarseVC": F, "-AER.BMD.FineVC": F, "-AER.BMD.RAOT": F,
    "-AER.BMD.CM.MRwa": F, "-AER.BMD.CM.MIwa": F, "-AER.BMD.CM.MRwaref": F, "-AER.BMD.CM.MIwaref": F, "-AER.BMD.CM.SDradius": F,
    "-AER.BMD.CM.SDvar": F, "-AER.BMD.FM.MRwa": F, "-AER.BMD.FM.MIwa": F, "-AER.BMD.FM.MRwaref": F, "-AER.BMD.FM.MIwaref": F,
    "-AER.BMD.FM.SDradius": F, "-AER.BMD.FM.SDvar": F, "-AER.ExtData": S, "-AER.DefMixture": S,
    "-AP.Log": S, "-AP.MOT": F, "-AP.HR": F, "-AP.AerProfile.Type": I, "-AP.AerHS.HA": F, "-AP.AerLayer.Zmin": F,
    "-AP.AerLayer.Zmax": F, "-AP.Psurf": F, "-AP.H2O": F, "-AP.O3": F, "-AP.CO2": F, "-AP.CH4": F, "-AP.AbsProfile.Type": I,
    "-AP.SpectralResol": F, "-AP.AbsProfile.UserFile": S,
    "-SOS.ResBin": S, "-SOS.ResFileUp": S, "-SOS.ResFileDown": S, "-SOS.ResFileUp.UserAng": S, "-SOS.ResFileDown.UserAng": S,
    "-SOS.Log": S, "-SOS.AbsModeCKD": I, "-SOS.Trans": S, "-SOS.Flux": S, "-SOS.OutputAlt": F, "-SOS.Ipolar": I, "-SOS.IGmax": I,
    "-SOS.View": I, "-SOS.View.Phi": F, "-SOS.View.Dphi": I,
}

# inc/SOS.h:85-87, 383, 508-535
DEFAULTS = {"-SOS.ResBin": "SOS_Result.bin", "-SOS.ResFileUp": "SOS_Up.txt", "-SOS.ResFileDown": "SOS_Down.txt", "-SOS.IGmax": 100,
            "-SOS.Ipolar": 1, "-SOS.OutputAlt": -1.0, "-AER.Tronca": 1}
REQUIRED = ("-SOS_Main.Wa", "-SOS_Main.ResRoot", "-ANG.Thetas", "-SURF.Type", "-SURF.Alb", "-AP.AerProfile.Type", "-AP.AbsProfile.Type",
            "-SOS.View")


def _number(txt):
    return float(txt.replace("D", "E").replace("d", "e"))          # the reference reads values list-directed: 1.D-3 is a number


def parse(argv, require=True):
    """`-KEYWORD value` pairs -> dict {keyword: value} with the optional keywords' defaults filled in.  Unknown keywords, missing
    values and (with require) missing required keywords raise ValueError, where the reference prints its usage text and stops."""
    argv = list(argv)
    if len(argv) % 2:
        raise ValueError("keywords and values come in pairs")
    out = dict(DEFAULTS)
    for k, v in zip(argv[0::2], argv[1::2]):
        if k not in KEYWORDS:
            raise ValueError("unknown keyword %s" % k)
        t = KEYWORDS[k]
        try:
            out[k] = str(v) if t == S else (int(_number(v)) if t == I else _number(v))
        except ValueError:
            raise ValueError("value %r of %s is not a number" % (v, k))
    if require:
        missing = [k for k in REQUIRED if k not in out]
        if missing:
            raise ValueError("missing required keyword(s): " + " ".join(missing))
    return out

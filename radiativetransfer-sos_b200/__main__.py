"""Command line of the device front end, with the keywords of SOS_ABS_MAIN.exe (exe/runSOS-ABS_demo.ksh):

    python -m radiativetransfer-sos_b200 -SOS_Main.Wa 0.910 -SOS_Main.ResRoot RESULT -ANG.Thetas 35. -SOS.View 1 ... [--wavelengths 0.865,0.910]

Exit status 0 on success, 1 on an error (message on stderr), as the reference's main program.  No CPU fallback: without a usable
CUDA device the run stops before computing anything."""
import sys


def main(argv):
    from . import api, frontend, keywords
    wavelengths = None
    if "--wavelengths" in argv:                                   # extension: a list of wavelengths in one run (not a reference keyword)
        i = argv.index("--wavelengths")
        wavelengths = [float(x) for x in argv[i + 1].split(",")]
        argv = argv[:i] + argv[i + 2:]
    try:
        kw = keywords.parse(argv)
    except ValueError as e:
        sys.stderr.write("  SOS_ABS_MAIN (device front end) : %s\n  known keywords: %s\n" % (e, " ".join(sorted(keywords.KEYWORDS))))
        return 1
    try:
        solver = api.Solver(0)
    except RuntimeError as e:
        sys.stderr.write("  SOS_ABS_MAIN (device front end) : %s\n" % e)
        return 1
    try:
        res, _ = frontend.run(solver, kw, wavelengths)
    except (ValueError, NotImplementedError, RuntimeError, OSError) as e:
        sys.stderr.write("  SOS_ABS_MAIN (device front end) : %s\n" % e)
        return 1
    finally:
        solver.close()
    for d in res.dirs:
        print("  results in", d)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

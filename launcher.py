"""``python launcher.py --num-process ... --exec-file main.py ...`` — same command line as the reference's launcher;
the implementation is ``adapcc_b200/launcher.py`` (torchrun instead of mpirun)."""
from adapcc_b200.launcher import main

if __name__ == "__main__":
    raise SystemExit(main())

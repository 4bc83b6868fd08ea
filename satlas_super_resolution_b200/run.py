"""python -m satlas_super_resolution_b200.run <module> [args...]   e.g.  ... run ssr.infer -opt ssr/options/infer_example.yml

Installs the drop-in (dropin.install) and then executes the reference's entry module UNCHANGED with runpy."""
import runpy
import sys

from . import dropin


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    module = sys.argv[1]
    sys.argv = [module] + sys.argv[2:]
    dropin.install(reference_root=".")
    runpy.run_module(module, run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()

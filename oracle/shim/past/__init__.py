"""Test-infrastructure stub for the `past` package (from `future`), which the
unmodified reference imports (pyGPs/Core/*.py: `from past.utils import old_div`).
Only used by oracle/gen_golden.py in the build container."""

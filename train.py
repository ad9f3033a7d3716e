"""`python train.py --model=planar --yaml=planar ...` — the reference's command line, served by marf_b200."""
from marf_b200.train import main

if __name__ == "__main__":
    main()

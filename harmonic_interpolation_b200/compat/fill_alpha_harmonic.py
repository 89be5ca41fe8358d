"""Top-level `fill_alpha_harmonic` (the reference's CLI name): put this directory on PYTHONPATH.

    python3 fill_alpha_harmonic.py path/to/input_image path/to/output_image
"""
from harmonic_interpolation_b200.fill_alpha_harmonic import *  # noqa: F401,F403
from harmonic_interpolation_b200.fill_alpha_harmonic import arr2img, fill_alpha_harmonic, load_img2arr, main  # noqa: F401

if __name__ == "__main__":
    main()

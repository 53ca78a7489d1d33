"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): tiny Marigold + GeoWizard pipelines."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import engine_checks as EC
from diffusion_e2e_ft_b200 import MarigoldPipeline
MarigoldPipeline.use_cuda_graph = False
print(EC.run_marigold_tiny())
print(EC.run_geowizard_tiny())

"""zpaq-v_b200: B200-native ZPAQ block codec behind the dy-tea/zpaq-v Compressor/Decompresser API.

csrc/      CUDA kernels (sm_100a) and the C ABI (include/zpaqgpu.h) -> libzpaqgpu.so
binding.py ctypes marshalling of the C ABI
codec.py   host mirror of the reference's Compressor / Decompresser / Reader / Writer
jidac.py   host mirror of the reference's JidacArchive (+ a reader)
vshim/     the V-side binding a maintainer of the reference would add (not compilable here)
"""
from . import binding
from .binding import Context, Multi, ZpaqGpuError, describe_model, level_header, tables
from .codec import Compressor, Decompresser, FileReader, FileWriter
from .jidac import JidacArchive
from . import jidac

__all__ = ["binding", "Context", "Multi", "ZpaqGpuError", "level_header", "tables", "describe_model", "Compressor", "Decompresser",
           "FileReader", "FileWriter", "JidacArchive", "jidac"]

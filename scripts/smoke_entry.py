"""Runs __graft_entry__.smoke() the way the driver does (one small checked invocation of the hot path on cuda:0)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import __graft_entry__ as g
g.smoke()

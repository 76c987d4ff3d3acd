"""Activates the engine-backed ntjoin_utils functions without editing ntJoin's bin/ directory.

Usage:  PYTHONPATH=<repo>/dropin:<repo>  NTJOIN_B200=1  ntJoin assemble ...
Python imports `sitecustomize` at start-up; this one registers an import hook that patches the
reference's `ntjoin_utils` (and `Ntjoin.print_graph` in `ntjoin`) right after they are loaded (see
ntjoin_b200/dropin.py).
"""
import importlib.abc
import importlib.util
import os
import sys

if os.environ.get("NTJOIN_B200", "0") not in ("", "0"):
    class _Patch(importlib.abc.MetaPathFinder):
        def find_spec(self, name, path, target=None):
            if name == "btllib":
                # bin/ntjoin_assemble.py:16 imports btllib for SeqReader (:313-316) and Indexlr (:478-481).  When the real
                # package is installed it is used as is; when it is not, the engine's btllib-shaped module serves both.
                sys.meta_path.remove(self)
                try:
                    real = importlib.util.find_spec(name)
                except (ImportError, ValueError):
                    real = None
                finally:
                    sys.meta_path.insert(0, self)
                if real is not None:
                    return None

                class _Alias(importlib.abc.Loader):
                    def create_module(self, spec):
                        import ntjoin_b200.btllib_compat as compat
                        return compat

                    def exec_module(self, module):
                        pass

                return importlib.util.spec_from_loader(name, _Alias())
            if name not in ("ntjoin_utils", "ntjoin", "ntjoin_assemble"):
                return None
            sys.meta_path.remove(self)
            try:
                spec = importlib.util.find_spec(name)
            finally:
                sys.meta_path.insert(0, self)
            if spec is None or spec.loader is None:
                return None
            loader = spec.loader
            orig_exec = loader.exec_module

            def exec_module(module):
                orig_exec(module)
                from ntjoin_b200 import dropin
                if name == "ntjoin_utils":
                    dropin.install(module)
                elif name == "ntjoin":
                    dropin.install_print_graph(module)      # Ntjoin.print_graph -> .mx.dot from arrays
                else:
                    dropin.install_scaffolder(module)       # NtjoinScaffolder.find_mx_min_max from arrays

            loader.exec_module = exec_module
            return spec

    sys.meta_path.insert(0, _Patch())

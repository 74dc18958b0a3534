"""The reference's flagship example (``/root/reference/examples/basic_example.py``) is not copied into this repo.
``tests/golden/basic_example_model.py`` holds the two classes of that file as a user would have them after
switching frameworks: the SAME class bodies, with the two framework imports changed
(``import jaxabm as jx`` -> ``import jaxabm_b200 as jx``, ``import jax.numpy as jnp`` -> ``import jaxabm_b200.numpy
as jnp``) and the plotting driver dropped.  When the reference tree is present (this container), ``check_against_
reference()`` verifies that claim by rewriting the reference file's imports in memory and comparing the class
sources; on the GPU box only the committed module is used."""
import ast
import os


REF = "/root/reference/examples/basic_example.py"
HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL = os.path.join(os.path.dirname(HERE), "golden", "basic_example_model.py")
REF_SENS = "/root/reference/examples/sensitivity/simple_sensitivity_example.py"
LOCAL_SENS = os.path.join(os.path.dirname(HERE), "golden", "simple_sensitivity_model.py")


def _class_sources(src: str):
    tree = ast.parse(src)
    return {n.name: ast.dump(n) for n in tree.body if isinstance(n, ast.ClassDef)}


def check_against_reference() -> bool:
    """True if verified, False if the reference tree is absent."""
    if not os.path.exists(REF):
        return False
    for ref_path, local in ((REF, LOCAL), (REF_SENS, LOCAL_SENS)):
        ref = open(ref_path).read().replace("import jaxabm as jx", "import jaxabm_b200 as jx") \
                                   .replace("import jax.numpy as jnp", "import jaxabm_b200.numpy as jnp")
        a, b = _class_sources(ref), _class_sources(open(local).read())
        assert set(b) <= set(a) and all(a[k] == b[k] for k in b), f"{local} drifted from the reference example"
    return True


def load(path: str):
    import importlib.util
    spec = importlib.util.spec_from_file_location(os.path.splitext(os.path.basename(path))[0], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

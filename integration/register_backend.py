#!/usr/bin/env python
"""Applies the backend registration of INTEGRATION.md §2 to the reference's files — at BUILD time, into a git-ignored
directory that shadows the originals on the include path (integration/_ref/gen/).  The reference tree is never modified and
no reference source is committed: this script only knows the ~10 anchor lines it inserts after, and fails loudly when an
anchor is missing (i.e. when the reference changes).

Edits (reference file:line of the anchor):
  include/plssvm/backend_types.hpp:30-43     enumerator `b200`
  include/plssvm/backend_types.hpp:66-72     forward declaration `namespace b200x { class csvm; }`
  include/plssvm/backend_types.hpp:93-149    `csvm_to_backend_type<b200x::csvm>`
  src/plssvm/backend_types.cpp:28-46         list_available_backends(): b200 if PLSSVM_HAS_B200_BACKEND
  src/plssvm/backend_types.cpp:48-71         determine_default_backend(): b200 first for gpu_nvidia
  src/plssvm/backend_types.cpp:73-112        operator<< / operator>> spelling "b200"
  include/plssvm/csvm_factory.hpp:123-140    include of the backend header + `case backend_type::b200`
"""
import os
import sys


def patch(text: str, anchor: str, insert: str, *, before: bool = False, path: str = "") -> str:
    if text.count(anchor) != 1:
        raise SystemExit(f"register_backend: anchor not found exactly once in {path}: {anchor!r} ({text.count(anchor)} matches)")
    return text.replace(anchor, insert + anchor if before else anchor + insert)


def main(ref: str, gen: str) -> None:
    # ---- backend_types.hpp ---------------------------------------------------------------------------------------------
    path = os.path.join(ref, "include/plssvm/backend_types.hpp")
    t = open(path).read()
    t = patch(t, "    sycl\n};", "", path=path).replace("    sycl\n};", "    sycl,\n    /** The Blackwell-native B200 backend (libplssvm_b200.so). */\n    b200\n};")
    t = patch(t, "namespace dpcpp { class csvm; }\n", "namespace b200x { class csvm; }\n", path=path)
    t = patch(t, "template <>\nstruct csvm_to_backend_type<cuda::csvm> {",
              "template <>\nstruct csvm_to_backend_type<b200x::csvm> {\n    /// The enum value representing the B200 backend.\n    static constexpr backend_type value = backend_type::b200;\n};\n",
              before=True, path=path)
    out = os.path.join(gen, "plssvm/backend_types.hpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    open(out, "w").write(t)

    # ---- backend_types.cpp ---------------------------------------------------------------------------------------------
    path = os.path.join(ref, "src/plssvm/backend_types.cpp")
    t = open(path).read()
    t = patch(t, "    return available_backends;\n", "#if defined(PLSSVM_HAS_B200_BACKEND)\n    available_backends.push_back(backend_type::b200);\n#endif\n", before=True, path=path)
    t = patch(t, "decision_order_type{ target_platform::gpu_nvidia, { backend_type::cuda,", "", path=path).replace(
        "decision_order_type{ target_platform::gpu_nvidia, { backend_type::cuda,", "decision_order_type{ target_platform::gpu_nvidia, { backend_type::b200, backend_type::cuda,")
    t = patch(t, "        case backend_type::sycl:\n            return out << \"sycl\";\n", "        case backend_type::b200:\n            return out << \"b200\";\n", path=path)
    t = patch(t, "    } else if (str == \"sycl\") {\n        backend = backend_type::sycl;\n", "    } else if (str == \"b200\") {\n        backend = backend_type::b200;\n", path=path)
    open(os.path.join(gen, "backend_types.cpp"), "w").write(t)

    # ---- csvm_factory.hpp ----------------------------------------------------------------------------------------------
    path = os.path.join(ref, "include/plssvm/csvm_factory.hpp")
    t = open(path).read()
    t = patch(t, "// only include requested/available backends\n", "#if defined(PLSSVM_HAS_B200_BACKEND)\n    #include \"b200_csvm.hpp\"  // plssvm::b200x::csvm (this repository: integration/b200_csvm.hpp)\n#endif\n", path=path)
    t = patch(t, "        case backend_type::sycl:\n            return make_csvm_sycl_impl(std::forward<Args>(args)...);\n",
              "#if defined(PLSSVM_HAS_B200_BACKEND)\n        case backend_type::b200:\n            return make_csvm_default_impl<b200x::csvm>(std::forward<Args>(args)...);\n#else\n        case backend_type::b200:\n            break;\n#endif\n",
              path=path)
    open(os.path.join(gen, "plssvm/csvm_factory.hpp"), "w").write(t)
    print(f"register_backend: wrote patched backend_types.hpp / backend_types.cpp / csvm_factory.hpp to {gen}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

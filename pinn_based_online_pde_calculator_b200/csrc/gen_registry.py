"""Emit jet_registry.cu from jet_configs.txt (list of kernel instantiations)."""
import sys

cfgs = [l.split() for l in open(sys.argv[1]) if l.strip() and not l.startswith("#")]
print('#include "jet_launch.h"')
for wp, n1, n2, mx in cfgs:
    print(f"extern const JetKernelInfo pinn_jet_info_{wp}_{n1}{n2}{mx};")
print("static const JetKernelInfo* const g_kernels[] = {")
for wp, n1, n2, mx in cfgs:
    print(f"  &pinn_jet_info_{wp}_{n1}{n2}{mx},")
print("};")
print("int pinn_kernel_count() { return (int)(sizeof(g_kernels) / sizeof(g_kernels[0])); }")
print("const JetKernelInfo* pinn_kernel_at(int i) { return g_kernels[i]; }")
print("const JetKernelInfo* pinn_find_kernel(int wp, int n1, int n2, int mix) {")
print("  for (int i = 0; i < pinn_kernel_count(); ++i) {")
print("    const JetKernelInfo* k = g_kernels[i];")
print("    if (k->wp == wp && k->n1 == n1 && k->n2 == n2 && k->mix == mix) return k;")
print("  }")
print("  return nullptr;")
print("}")

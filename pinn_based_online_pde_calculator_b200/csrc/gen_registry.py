"""Emit jet_registry.cu from jet_configs.txt / jet_mma_configs.txt (kernel instantiations)."""
import sys

def read(path):
    return [l.split() for l in open(path) if l.strip() and not l.startswith("#")]

simt, mma = read(sys.argv[1]), read(sys.argv[2])
tc = read(sys.argv[3]) if len(sys.argv) > 3 else []
print('#include "jet_launch.h"')
for wp, n1, n2, mx in simt:
    print(f"extern const JetKernelInfo pinn_jet_info_{wp}_{n1}{n2}{mx};")
for wp, n1, n2, mx in mma:
    print(f"extern const JetKernelInfo pinn_mma_info_{wp}_{n1}{n2}{mx};")
for wp, n1, n2, mx in tc:
    print(f"extern const JetKernelInfo pinn_tc_info_{wp}_{n1}{n2}{mx};")
print("static const JetKernelInfo* const g_kernels[] = {")
for wp, n1, n2, mx in simt:
    print(f"  &pinn_jet_info_{wp}_{n1}{n2}{mx},")
for wp, n1, n2, mx in mma:
    print(f"  &pinn_mma_info_{wp}_{n1}{n2}{mx},")
for wp, n1, n2, mx in tc:
    print(f"  &pinn_tc_info_{wp}_{n1}{n2}{mx},")
print("};")
print("int pinn_kernel_count() { return (int)(sizeof(g_kernels) / sizeof(g_kernels[0])); }")
print("const JetKernelInfo* pinn_kernel_at(int i) { return g_kernels[i]; }")
print("const JetKernelInfo* pinn_find_kernel(int wp, int n1, int n2, int mix, int kind) {")
print("  for (int i = 0; i < pinn_kernel_count(); ++i) {")
print("    const JetKernelInfo* k = g_kernels[i];")
print("    if (k->wp == wp && k->n1 == n1 && k->n2 == n2 && k->mix == mix && k->kind == kind) return k;")
print("  }")
print("  return nullptr;")
print("}")

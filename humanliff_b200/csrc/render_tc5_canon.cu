// Second compilation of render_tc5.cu: the kernel of the canonical-mode launches (k_render_tc5_canon, compact MLP copy) and its
// launcher hl_r5_launch_canon -- see the comment above the kernel in render_tc5.cu.
#define HL_R5_CANON_TU 1
#include "render_tc5.cu"

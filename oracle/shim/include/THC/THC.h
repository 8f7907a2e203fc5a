/* Stand-in for the THC umbrella header that torch removed; the names the reference's nms_kernel.cu takes from
 * it are provided by oracle/shim/ref_compat_nms.h.  Test infrastructure only. */
#pragma once

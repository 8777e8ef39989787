/* Second C translation unit of the K0 example: it includes swgl.h too, which the reference's own
 * header cannot survive (swgl.h:19-20 define objects in the header; two C units fail to link). */
#include "swgl.h"

/* the survey's hash: h = (h ^ word) * 1099511628211, seed 1469598103934665603 */
unsigned long long k0_fnv1a64(const uint32_t* words, unsigned long long n)
{
    unsigned long long h = 1469598103934665603ull;
    for (unsigned long long i = 0; i < n; i++) h = (h ^ words[i]) * 1099511628211ull;
    return h;
}

/* TEST INFRASTRUCTURE ONLY -- force-included (-include) when compiling the
 * unmodified reference sources from /root/reference into oracle/_ref.
 * It forward-declares the structs that client/render.h:44-53 and
 * common/mathlib.h:89 only mention inside prototypes (gcc would otherwise
 * give them prototype scope). Nothing here changes reference behaviour. */
struct mpic_s; struct cplane_s; struct model_s; struct edict_s;

/* oracle shim (test infrastructure): only the enum name that include/utils.h:3,20 mentions. */
#ifndef ORACLE_SHIM_ASSIMP_MATERIAL_H
#define ORACLE_SHIM_ASSIMP_MATERIAL_H
enum aiTextureType { aiTextureType_NONE = 0 };
#endif

// Host-only check of b2r_tileset_write_exr / _write_tiff (test infrastructure): builds a tile
// set of STORED tiles by hand -- level (lv), pixel (x, y), channel c holds
// ((7x + 13y + 17c + 29lv) mod 1000) / 1000 - 0.25 -- and writes it; the Python
// side (tests/test_cpu.py::test_exr_and_tiff_writers_on_host) reads the file back with
// tests/exr_mini.py and with the OpenEXR library.  Includes the product
// translation unit to see the tile-set structure; linked against the other
// objects of the library build (openimageio_b200/csrc/build/*.o).
// usage: exr_harness out.exr|out.tif W H nch half(0|1) tile [bigtiff]
#include "b2r_api_tiles.cpp"
#include <cuda_fp16.h>
static float val(int lv,int y,int x,int c){ return float((x*7+y*13+c*17+lv*29)%1000)/1000.0f - 0.25f; }
int main(int argc,char**argv){
  const char* path=argv[1]; int W=atoi(argv[2]),H=atoi(argv[3]),nch=atoi(argv[4]),half=atoi(argv[5]),tile=atoi(argv[6]);
  b2r_tileset ts; ts.nch=nch; ts.tile_w=ts.tile_h=tile; ts.out_basetype=half?B2R_HALF:B2R_FLOAT; ts.compression=0; ts.predictor=0; ts.first_level=0;
  int w=W,h=H,lv=0; const int es=half?2:4;
  for(;;){
    b2r_tileset::Level L; L.w=w;L.h=h;L.ntx=(w+tile-1)/tile;L.nty=(h+tile-1)/tile; L.tiles.resize(size_t(L.ntx)*L.nty);
    for(int ty=0;ty<L.nty;++ty)for(int tx=0;tx<L.ntx;++tx){
      auto&t=L.tiles[size_t(ty)*L.ntx+tx]; t.assign(size_t(tile)*tile*nch*es,0);
      for(int yy=0;yy<tile;++yy)for(int xx=0;xx<tile;++xx){int y=ty*tile+yy,x=tx*tile+xx; if(y>=h||x>=w)continue;
        for(int c=0;c<nch;++c){ float v=val(lv,y,x,c); unsigned char*p=t.data()+((size_t(yy)*tile+xx)*nch+c)*es;
          if(half){ __half hv=__float2half(v); memcpy(p,&hv,2);} else memcpy(p,&v,4);} }
    }
    ts.levels.push_back(std::move(L));
    if(w==1&&h==1)break; w=std::max(1,w/2);h=std::max(1,h/2);++lv;
  }
  // a third attribute with a name beyond 31 characters when asked for ("long"): the
  // writer must then set the long-names version flag
  const char* names[3]={"textureformat","oiio:SHA-1","an_attribute_name_of_more_than_31_characters"};
  const char* vals[3]={"Plain Texture","0123ABCD","x"};
  b2r_exr_options o; memset(&o,0,sizeof(o)); o.attr_names=names;o.attr_values=vals;
  o.nattrs=(argc>7&&!strcmp(argv[7],"long"))?3:2;o.nthreads=4;
  char err[256]="";
  const size_t pl=strlen(path);
  int rc;
  if(pl>4&&!strcmp(path+pl-4,".tif")){  // the TIFF container around the same stored tiles
    b2r_tiff_options fo; memset(&fo,0,sizeof(fo)); fo.image_description="oiio:SHA-1=0123ABCD"; fo.bigtiff=argc>7?atoi(argv[7]):0;
    rc=b2r_tileset_write_tiff(&ts,path,&fo,err,sizeof(err));
  } else
    rc=b2r_tileset_write_exr(&ts,path,&o,err,sizeof(err));
  printf("rc=%d %s levels=%d\n",rc,err,(int)ts.levels.size());
  return rc;
}
